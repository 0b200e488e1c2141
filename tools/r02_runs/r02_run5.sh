#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_blockmix_gpu.py -q -x -k "smalln" > $O/r02_pytest_smalln.log 2>&1; echo "smalln rc=$?"; tail -40 $O/r02_pytest_smalln.log
timeout 900 python -m pytest tests -m gpu -q > $O/r02_pytest_gpu5.log 2>&1; echo "pytest rc=$?"; tail -15 $O/r02_pytest_gpu5.log
timeout 300 python tools/bench_configs.py 2>&1 | head -3
timeout 300 python tools/stress.py dit64 4000 2>&1 | tail -2
