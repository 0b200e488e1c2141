#!/bin/bash
# usage: tools/sweep_lags.sh  (on the GPU box) - sweeps the schedule lags of the fused blockmix kernel
for cfg in "1 3" "2 5" "3 6" "4 8"; do
  set -- $cfg
  for nn in "" "--no-normalize"; do
    r=$(MHLA_LAG2=$1 MHLA_LAG3=$2 timeout 120 python bench.py --no-cpu-baseline --steps 30 --e2e-steps 1 $nn 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,1), round(d['roofline']['frac'],3))" 2>&1)
    echo "lag2=$1 lag3=$2 $nn : us/step, frac = $r"
  done
done
