#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
run() { local label=$1; shift
  env "$@" timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 1 --no-graph 2>/dev/null | python -c "import sys,json; [print('$label', round(json.loads(l)['ms_per_step']*1e3,2), round(json.loads(l)['roofline']['frac'],4)) for l in sys.stdin if l.startswith('{')]"
}
for i in 1 2; do
run v0fix MHLA_B200_LIB=$PWD/mhla_b200/libmhla_b200_v0fix.so
run cur X=1
run diag MHLA_B200_LIB=$PWD/mhla_b200/libmhla_b200_diag.so
done 2>&1 | tee $O/r02_ab8.log
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
timeout 100 python tools/stress.py selftest 2>&1 | tail -2
timeout 200 python tools/stress.py wan_norm 3000 2>&1 | tail -1
MHLA_STRESS_PRODUCT=1 timeout 200 python tools/stress.py wan_norm 3000 2>&1 | tail -1
