#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/r02_pytest_gpu11.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r02_pytest_gpu11.log
for i in 1 2; do
  for lib in v0 cur; do
    if [ $lib = v0 ]; then export MHLA_B200_LIB=$PWD/mhla_b200/libmhla_b200_v0.so; else unset MHLA_B200_LIB; fi
    timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "import sys,json; [print('$lib graph', json.loads(l)['ms_per_step'], json.loads(l)['roofline']['frac']) for l in sys.stdin if l.startswith('{')]"
    timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 1 --no-graph 2>/dev/null | python -c "import sys,json; [print('$lib eager', json.loads(l)['ms_per_step'], json.loads(l)['roofline']['frac']) for l in sys.stdin if l.startswith('{')]"
  done
done 2>&1 | tee $O/r02_ab_v0_cur2.log
unset MHLA_B200_LIB
timeout 600 python tools/bench_configs.py 2>/dev/null | tee $O/r02_configs_final.jsonl
