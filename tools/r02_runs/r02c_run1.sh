#!/bin/bash
# round 2, third session, run 1: A/B of the early accumulator hand-back variants (MHLA_EARLY_P1 / _P2 / _P3) + parity of the most aggressive ones
set -u
O=gpurun_out; mkdir -p $O
L=mhla_b200
timeout 600 python tools/ab_libs.py --reps 3 --shapes headline,nonorm,wan,wan_norm base=$L/libmhla_b200.so e1=$L/libmhla_b200_e1.so e3a=$L/libmhla_b200_e3a.so e3b=$L/libmhla_b200_e3b.so e13a=$L/libmhla_b200_e13a.so e13b=$L/libmhla_b200_e13b.so e123a=$L/libmhla_b200_e123a.so e123b=$L/libmhla_b200_e123b.so > $O/r02c_ab_early.log 2>&1
cat $O/r02c_ab_early.log | cut -c1-200
for v in e123a e123b; do
MHLA_B200_LIB=$PWD/$L/libmhla_b200_$v.so timeout 600 python -m pytest tests/test_blockmix_gpu.py tests/test_backward_gpu.py -m gpu -q -x 2>&1 | tail -2
done
