#!/bin/bash
# Final validation of the round: what the driver runs (smoke, -m gpu tests, both bench arms) + stress.
set -u
O=gpurun_out; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 900 python -m pytest tests -m gpu -q > $O/r02_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02_pytest_gpu_final.log
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r02_bench_reference_final.json 2>&1; tail -c 400 $O/r02_bench_reference_final.json; echo
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02_bench_final.json 2> $O/r02_bench_final.err; python -c "
import json
r=json.loads([x for x in open('$O/r02_bench_final.json') if x.startswith('{')][-1])
print('ms', r['ms_per_step'], 'frac', r['roofline']['frac'], 'traffic', r['roofline']['traffic'], 'e2e ms', r['e2e']['ms_per_step'], 'cpu', r['cpu_baseline']['value'], r['clocks'])"
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-normalize --no-cpu-baseline > $O/r02_bench_final_nonorm.json 2>/dev/null; python -c "
import json
r=json.loads([x for x in open('$O/r02_bench_final_nonorm.json') if x.startswith('{')][-1]); print('nonorm ms', r['ms_per_step'], 'frac', r['roofline']['frac'])"
timeout 600 python tools/bench_configs.py 2>/dev/null | tee $O/r02_configs_final2.jsonl | cut -c1-120
L=$O/r02_stress_product_final.log; : > $L
for v in "wan_norm 6000" "rn_d64 6000" "n_d128 6000" "headline 6000" "dit64 6000" "wan 4000"; do
  echo "--- $v" >> $L; MHLA_STRESS_PRODUCT=1 timeout 300 python tools/stress.py $v 2>&1 | tail -1 >> $L
done; cat $L
