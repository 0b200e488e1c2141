#!/bin/bash
# round 2, third session: what the driver runs (smoke, -m gpu tests, both bench arms) + configs + Wan layer + stress on the final build
set -u
O=gpurun_out; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -q > $O/r02c_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02c_pytest_gpu_final.log
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r02c_bench_reference_final.json 2>&1; tail -c 300 $O/r02c_bench_reference_final.json; echo
show() { python -c "
import json,sys
r=json.loads([x for x in open('$1') if x.startswith('{')][-1])
print('$1', 'ms', round(r['ms_per_step'],5), 'frac', round(r['roofline']['frac'],4), 'e2e ms', round(r['e2e']['ms_per_step'],3), r['clocks'])"; }
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02c_bench_final.json 2> $O/r02c_bench_final.err; show $O/r02c_bench_final.json
for c in 16 32; do timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --e2e-chunks $c > $O/r02c_bench_chunks$c.json 2>/dev/null; show $O/r02c_bench_chunks$c.json; done
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-normalize --no-cpu-baseline > $O/r02c_bench_final_nonorm.json 2>/dev/null; show $O/r02c_bench_final_nonorm.json
timeout 600 python tools/bench_configs.py 2>/dev/null | tee $O/r02c_configs.jsonl | cut -c1-120
timeout 300 python tools/wan_layer_bench.py > $O/r02c_wan_layer.log 2>&1; tail -4 $O/r02c_wan_layer.log
L=$O/r02c_stress_product_final.log; : > $L
for v in "wan_norm 4000" "rn_d64 4000" "headline 6000" "dit64 4000" "wan 4000"; do
  echo "--- $v" >> $L; MHLA_STRESS_PRODUCT=1 timeout 300 python tools/stress.py $v 2>&1 | tail -1 >> $L
done; cat $L
