#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_modules.py -m gpu -q > $O/r02_pytest_wan2.log 2>&1; echo "modules rc=$?"; tail -4 $O/r02_pytest_wan2.log
timeout 300 python tools/wan_layer_bench.py 2>&1 | tee $O/r02_wan_layer.log
timeout 900 python bench.py --impl reference-gpu --sweep > $O/r02_comparators_stdout.json 2> $O/r02_comparators.err; grep '"N"' $O/r02_comparators.err | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['N'], {k:round(v,1) for k,v in r.items() if k.endswith('_us') or k.startswith('speedup')})"
