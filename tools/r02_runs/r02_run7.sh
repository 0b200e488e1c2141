#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_blockmix_gpu.py -q -x -k "graph or smalln or vs_oracle" > $O/r02_pytest_smalln2.log 2>&1; echo "rc=$?"; tail -5 $O/r02_pytest_smalln2.log
for b in 2 64 256; do timeout 120 python tools/prof_smalln.py $b; done 2>&1 | tee $O/r02_prof_smalln.log
timeout 600 python tools/bench_configs.py 2>/dev/null | head -5
