#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
run() { local label=$1; shift
  env "$@" timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 1 --no-graph 2>/dev/null | python -c "import sys,json; [print('$label', round(json.loads(l)['ms_per_step']*1e3,2), round(json.loads(l)['roofline']['frac'],4)) for l in sys.stdin if l.startswith('{')]"
}
for i in 1 2; do
run v0 MHLA_B200_LIB=$PWD/mhla_b200/libmhla_b200_v0.so
run v0fix MHLA_B200_LIB=$PWD/mhla_b200/libmhla_b200_v0fix.so
run cur X=1
run cur_nonorm_skip X=1
done 2>&1 | tee $O/r02_ab4.log
timeout 200 python tools/stress.py wan_norm 3000 2>&1 | tail -1
MHLA_B200_LIB=$PWD/mhla_b200/libmhla_b200_v0fix.so timeout 100 python tools/prof_roles.py > $O/r02_prof_roles_v0fix.log 2>&1
timeout 100 python tools/prof_roles.py > $O/r02_prof_roles_cur.log 2>&1
paste $O/r02_prof_roles_v0fix.log $O/r02_prof_roles_cur.log | cut -c1-250
