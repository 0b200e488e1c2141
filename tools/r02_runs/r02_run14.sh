#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
run() { local label=$1; shift
  env "$@" timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 1 --no-graph $EXTRA 2>/dev/null | python -c "import sys,json; [print('$label', round(json.loads(l)['ms_per_step']*1e3,2), round(json.loads(l)['roofline']['frac'],4)) for l in sys.stdin if l.startswith('{')]"
}
EXTRA=""
for i in 1 2; do
run v0 MHLA_B200_LIB=$PWD/mhla_b200/libmhla_b200_v0.so
run v0fix MHLA_B200_LIB=$PWD/mhla_b200/libmhla_b200_v0fix.so
run cur X=1
done 2>&1 | tee $O/r02_ab5.log
EXTRA="--no-normalize"; run cur_nonorm X=1 | tee -a $O/r02_ab5.log
timeout 900 python -m pytest tests -m gpu -q -x > $O/r02_pytest_gpu14.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r02_pytest_gpu14.log
L=$O/r02_stress_final.log; : > $L
for v in "wan_norm 6000" "rn_d64 6000" "n_d128 6000" "rn_w256 6000" "rn_b1 6000" "wan 4000" "headline 6000" "dit64 6000" "small_rope 6000"; do
  echo "--- $v" >> $L
  timeout 300 python tools/stress.py $v 2>&1 | grep -v "^Search\|^CUDA kernel\|^For debugging\|^Compile with\|^$" >> $L
done
grep -c "0 sampled mismatches" $L; grep -i "fail" $L
