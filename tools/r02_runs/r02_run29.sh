#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
for k in 0 1; do
echo "MIX_HI_ONLY=$k"
MHLA_MIX_HI_ONLY=$k timeout 300 python tools/bench_configs.py 2>/dev/null | grep -E "cfg4|headline"
MHLA_MIX_HI_ONLY=$k MHLA_STRESS_PRODUCT=1 timeout 200 python tools/stress.py wan 200 2>&1 | head -1
MHLA_MIX_HI_ONLY=$k MHLA_STRESS_PRODUCT=1 timeout 200 python tools/stress.py wan_norm 200 2>&1 | head -1
done 2>&1 | tee $O/r02_mix_hi_only.log
