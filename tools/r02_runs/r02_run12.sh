#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
run() { # label, env...
  local label=$1; shift
  env "$@" timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 1 --no-graph 2>/dev/null | python -c "import sys,json; [print('$label', round(json.loads(l)['ms_per_step']*1e3,2), round(json.loads(l)['roofline']['frac'],4)) for l in sys.stdin if l.startswith('{')]"
}
for i in 1 2; do
run v0 MHLA_B200_LIB=$PWD/mhla_b200/libmhla_b200_v0.so
run cur X=1
run noqwait MHLA_B200_LIB=$PWD/mhla_b200/libmhla_b200_noqwait.so
run cur_p2tma0 MHLA_P2TMA=0
run cur_slots1 MHLA_SLOTS=1
run cur_ohint0 MHLA_OHINT=0
run cur_ra3 MHLA_RUNAHEAD=3
run cur_ra1 MHLA_RUNAHEAD=1
done 2>&1 | tee $O/r02_ab3.log
