#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/r02_pytest_gpu10.log 2>&1; echo "pytest rc=$?"; tail -25 $O/r02_pytest_gpu10.log
for i in 1 2 3; do
  for lib in v0 cur; do
    if [ $lib = v0 ]; then export MHLA_B200_LIB=$PWD/mhla_b200/libmhla_b200_v0.so; else unset MHLA_B200_LIB; fi
    timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "import sys,json; [print('$lib', json.loads(l)['ms_per_step'], json.loads(l)['roofline']['frac']) for l in sys.stdin if l.startswith('{')]"
  done
done 2>&1 | tee $O/r02_ab_v0_cur.log
