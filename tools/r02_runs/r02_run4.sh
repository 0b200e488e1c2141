#!/bin/bash
# GPU call 4: verify the parity-alias fix (stress on every preset that failed), then A/B the P2 store path.
set -u
O=gpurun_out; mkdir -p $O; L=$O/r02_stress_fixed.log; : > $L
for v in "wan_norm 6000" "wan_norm 6000" "rn_d64 6000" "n_d128 6000" "rn_w256 6000" "rn_b1 6000" "wan 4000" "headline 6000" "dit64 6000" "small_rope 6000"; do
  echo "--- $v" >> $L
  timeout 300 python tools/stress.py $v 2>&1 | grep -v "^Search\|^CUDA kernel\|^For debugging\|^Compile with\|^$" >> $L
done
echo "--- MHLA_P2TMA=1 wan_norm 6000" >> $L
MHLA_P2TMA=1 timeout 300 python tools/stress.py wan_norm 6000 2>&1 | grep -v "^Search\|^CUDA kernel\|^For debugging\|^Compile with\|^$" >> $L
cat $L
echo "== perf"
for knob in "MHLA_P2TMA=0" "MHLA_P2TMA=1" "MHLA_P2TMA=0 MHLA_WSHINT=0" ; do
  echo "$knob" >> $O/r02_perf4.log
  env $knob timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 1 2>&1 | python -c "import sys,json; [print(json.loads(l)['ms_per_step'], json.loads(l)['roofline']['frac']) for l in sys.stdin if l.startswith('{')]" >> $O/r02_perf4.log 2>&1
  env $knob timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 1 --no-normalize 2>&1 | python -c "import sys,json; [print('nonorm', json.loads(l)['ms_per_step'], json.loads(l)['roofline']['frac']) for l in sys.stdin if l.startswith('{')]" >> $O/r02_perf4.log 2>&1
done
cat $O/r02_perf4.log
timeout 600 python tools/bench_configs.py > $O/r02_configs_v3.jsonl 2>&1; cat $O/r02_configs_v3.jsonl
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q > $O/r02_pytest_gpu4.log 2>&1; echo "pytest rc=$?"; tail -8 $O/r02_pytest_gpu4.log
