#!/bin/bash
# round 2, third session, run 2: A/B of static first item / single-warp accumulator poll / low-register P3 early release on
# top of the new default (P1 early + P3 release after the last TMEM read); operator-level timing of the fused gate/add epilogue
set -u
O=gpurun_out; mkdir -p $O
L=mhla_b200
timeout 600 python tools/ab_libs.py --reps 3 --shapes headline,nonorm,wan,n8k base=$L/libmhla_b200.so old=$L/libmhla_b200_old.so sf=$L/libmhla_b200_sf.so p3lr=$L/libmhla_b200_p3lr.so poll1=$L/libmhla_b200_poll1.so sfpoll1=$L/libmhla_b200_sfpoll1.so > $O/r02c_ab_2.log 2>&1
cat $O/r02c_ab_2.log | cut -c1-200
for v in sfpoll1 p3lr; do
MHLA_B200_LIB=$PWD/$L/libmhla_b200_$v.so timeout 600 python -m pytest tests/test_blockmix_gpu.py tests/test_backward_gpu.py -m gpu -q -x 2>&1 | tail -2
done
timeout 600 python -m pytest tests/test_blockmix_gpu.py tests/test_backward_gpu.py tests/test_modules.py -m gpu -q -x 2>&1 | tail -2
timeout 300 python tools/wan_layer_bench.py 2>&1 | tail -3
