#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
run() { local label=$1; shift
  env "$@" timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 1 --no-graph $EXTRA 2>/dev/null | python -c "import sys,json; [print('$label', round(json.loads(l)['ms_per_step']*1e3,2), round(json.loads(l)['roofline']['frac'],4)) for l in sys.stdin if l.startswith('{')]"
}
for i in 1 2; do
EXTRA=""
run base MHLA_CLAIM_AHEAD=0
run ahead MHLA_CLAIM_AHEAD=1
run ahead_ra3 MHLA_CLAIM_AHEAD=1 MHLA_RUNAHEAD=3
EXTRA="--no-normalize"
run base_nonorm MHLA_CLAIM_AHEAD=0
run ahead_nonorm MHLA_CLAIM_AHEAD=1
done 2>&1 | tee $O/r02_ab_claim_ahead.log
MHLA_CLAIM_AHEAD=1 timeout 600 python -m pytest tests/test_blockmix_gpu.py -m gpu -q -x 2>&1 | tail -2
MHLA_CLAIM_AHEAD=1 MHLA_STRESS_PRODUCT=1 timeout 200 python tools/stress.py wan_norm 3000 2>&1 | tail -1
MHLA_CLAIM_AHEAD=1 MHLA_STRESS_PRODUCT=1 timeout 200 python tools/stress.py headline 3000 2>&1 | tail -1
