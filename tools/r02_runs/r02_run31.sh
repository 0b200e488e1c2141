#!/bin/bash
set -u
run() { local label=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "import sys,json; [print('$label', round(json.loads(l)['ms_per_step']*1e3,2), round(json.loads(l)['roofline']['frac'],4), json.loads(l)['clocks']) for l in sys.stdin if l.startswith('{')]"
}
for i in 1 2; do
run sample0.5ms X=1
run sample5ms MHLA_BENCH_SAMPLE_S=0.005
run sample50ms MHLA_BENCH_SAMPLE_S=0.05
done
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "import sys,json; [print('steps200', round(json.loads(l)['ms_per_step']*1e3,2), round(json.loads(l)['roofline']['frac'],4), json.loads(l)['clocks']) for l in sys.stdin if l.startswith('{')]"
