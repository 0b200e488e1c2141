#!/bin/bash
# round 2, third session: final validation after the wan_prep / LePE kernels - smoke, full -m gpu suite, bench, configs, short stress
set -u
O=gpurun_out; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -q > $O/r02c_pytest_gpu_final3.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02c_pytest_gpu_final3.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02c_bench_final3.json 2> $O/r02c_bench_final3.err
python -c "
import json
r=json.loads([x for x in open('$O/r02c_bench_final3.json') if x.startswith('{')][-1])
print('ms', round(r['ms_per_step'],5), 'frac', round(r['roofline']['frac'],4), 'traffic', r['roofline']['traffic'], 'e2e ms', round(r['e2e']['ms_per_step'],3), 'cpu', round(r['cpu_baseline']['value']), r['clocks'], 'launches', r['gpu_launches'])"
timeout 300 python tools/bench_configs.py 2>/dev/null | tee $O/r02c_configs2.jsonl | cut -c1-110
for v in "headline 3000" "wan_norm 2000" "dit64 2000"; do MHLA_STRESS_PRODUCT=1 timeout 200 python tools/stress.py $v 2>&1 | tail -1; done
