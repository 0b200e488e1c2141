#!/bin/bash
# GPU call 3: isolate the rope+normaliser launch failure (diagnostics self-test, shape / path variations, memcheck).
set -u
O=gpurun_out; mkdir -p $O; L=$O/r02_isolate.log; : > $L
timeout 120 python tools/stress.py selftest >> $L 2>&1
for v in "wan_norm 3000" "wan_norm 3000" "wan_norm 3000 three_launch" "rn_d64 3000" "n_d128 3000" "r_d128 3000" "rn_w256 3000" "rn_w128 3000" "rn_m128 3000" "rn_m64 3000" "rn_b1 3000" "small_rope 6000"; do
  echo "--- $v" >> $L
  timeout 200 python tools/stress.py $v 2>&1 | grep -v "^Search\|^CUDA kernel\|^For debugging\|^Compile with\|^$" >> $L
done
echo "--- MHLA_NO_SELF_PREP wan_norm" >> $L
MHLA_NO_SELF_PREP=1 timeout 200 python tools/stress.py wan_norm 3000 2>&1 | grep -v "^Search\|^CUDA kernel\|^For debugging\|^Compile with\|^$" >> $L
echo "--- v1 lib wan_norm" >> $L
MHLA_B200_LIB=$PWD/mhla_b200/libmhla_b200_v1.so timeout 200 python tools/stress.py wan_norm 3000 2>&1 | grep -v "^Search\|^CUDA kernel\|^For debugging\|^Compile with\|^$" >> $L
echo "--- v0 lib wan_norm (force_fused not available through the new shim: C-level three-launch)" >> $L
cat $L
echo "== memcheck 600 calls"
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/stress.py wan_norm 600 > $O/r02_memcheck_wan600.log 2>&1; echo "memcheck rc=$?"; grep -v "^=========     " $O/r02_memcheck_wan600.log | tail -12
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q > $O/r02_pytest_gpu3.log 2>&1; echo "pytest rc=$?"; tail -8 $O/r02_pytest_gpu3.log
