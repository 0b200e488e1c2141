#!/bin/bash
for cfg in "0 0" "1 0" "0 1" "1 1"; do
  set -- $cfg
  for nn in "" "--no-normalize"; do
    r=$(MHLA_DEPMODE=$1 MHLA_SIGMODE=$2 MHLA_LAG2=2 MHLA_LAG3=5 timeout 120 python bench.py --no-cpu-baseline --steps 30 --e2e-steps 1 $nn 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,1), round(d['roofline']['frac'],3))" 2>&1)
    echo "dep_mode=$1 sig_mode=$2 lag 2,5 $nn : us/step, frac = $r"
  done
done
