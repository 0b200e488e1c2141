#!/bin/bash
# A/B builds of the kernel library (tuning aid): benches each libvariant_*.so and the current library
for lib in mhla_b200/libvariant_A.so mhla_b200/libvariant_B.so mhla_b200/libvariant_D.so mhla_b200/libmhla_b200.so; do
  for nn in "" "--no-normalize"; do
    r=$(MHLA_B200_LIB=$PWD/$lib MHLA_LAG2=2 MHLA_LAG3=5 timeout 120 python bench.py --no-cpu-baseline --steps 30 --e2e-steps 1 $nn 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,1), round(d['roofline']['frac'],3))" 2>&1)
    echo "$lib $nn : us/step, frac = $r"
  done
done
