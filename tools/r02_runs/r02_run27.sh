#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 300 python tools/small_units.py 2>&1 | tee $O/r02_small_units.log
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r02_bench_1gpu_b.json 2>$O/r02_bench_1gpu_b.err; python -c "
import json
r=json.loads([x for x in open('$O/r02_bench_1gpu_b.json') if x.startswith('{')][-1]); print(r['ms_per_step'], r['roofline']['frac'], r['config']['launch'], r['e2e']['ms_per_step'], r.get('cpu_baseline'))"
