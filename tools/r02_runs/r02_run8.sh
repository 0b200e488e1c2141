#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_blockmix_gpu.py -q -x -k "graph or smalln" > $O/r02_pytest_smalln3.log 2>&1; echo "rc=$?"; tail -4 $O/r02_pytest_smalln3.log
for b in 64 256; do timeout 120 python tools/prof_smalln.py $b; done 2>&1 | tee $O/r02_prof_smalln2.log
timeout 600 python tools/bench_configs.py 2>/dev/null | head -5 | tee $O/r02_configs_smalln2.jsonl
timeout 300 python tools/causal_phases.py 2>&1 | tee $O/r02_causal_phases.log
