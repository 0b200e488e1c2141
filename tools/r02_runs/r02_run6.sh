#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_blockmix_gpu.py -q -x -k "graph or smalln" > $O/r02_pytest_graph.log 2>&1; echo "graph rc=$?"; tail -15 $O/r02_pytest_graph.log
timeout 600 python tools/bench_configs.py > $O/r02_configs_graph.jsonl 2>$O/r02_configs_graph.err; cat $O/r02_configs_graph.jsonl; tail -3 $O/r02_configs_graph.err
timeout 120 python tools/host_overhead.py
