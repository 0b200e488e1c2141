#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_blockmix_gpu.py -q -x -k "3d_block" > $O/r02_pytest_3d.log 2>&1; echo "3d rc=$?"; tail -30 $O/r02_pytest_3d.log
timeout 900 python -m pytest tests -m gpu -q > $O/r02_pytest_gpu19.log 2>&1; echo "all rc=$?"; tail -5 $O/r02_pytest_gpu19.log
for i in 1 2; do
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 1 --no-graph 2>/dev/null | python -c "import sys,json; [print('cur', round(json.loads(l)['ms_per_step']*1e3,2), round(json.loads(l)['roofline']['frac'],4)) for l in sys.stdin if l.startswith('{')]"
MHLA_B200_LIB=$PWD/mhla_b200/libmhla_b200_v0fix.so timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 1 --no-graph 2>/dev/null | python -c "import sys,json; [print('v0fix', round(json.loads(l)['ms_per_step']*1e3,2), round(json.loads(l)['roofline']['frac'],4)) for l in sys.stdin if l.startswith('{')]"
done
