#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
run() { local label=$1; shift
  env "$@" timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 1 --no-graph 2>/dev/null | python -c "import sys,json; [print('$label', round(json.loads(l)['ms_per_step']*1e3,2), round(json.loads(l)['roofline']['frac'],4)) for l in sys.stdin if l.startswith('{')]"
}
for i in 1 2; do
run v0fix MHLA_B200_LIB=$PWD/mhla_b200/libmhla_b200_v0fix.so
run cur_hint200us X=1
run hint20us MHLA_B200_LIB=$PWD/mhla_b200/libmhla_b200_exp_hint20000u.so
run hint2ms MHLA_B200_LIB=$PWD/mhla_b200/libmhla_b200_exp_hint2000000u.so
done 2>&1 | tee $O/r02_ab7.log
timeout 300 python -m pytest tests/test_blockmix_gpu.py -m gpu -q -x 2>&1 | tail -2
