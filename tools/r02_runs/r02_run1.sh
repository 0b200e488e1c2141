#!/bin/bash
# GPU call 1 of round 2: tests, stall hunt (stress + compute-sanitizer), A/B of the round-1 kernel (v0) against the
# TMA-store / lazy-signal kernel (v1), per-role profile, comparators.  Everything lands in gpurun_out/.
set -u
O=gpurun_out; mkdir -p $O
V0=$PWD/mhla_b200/libmhla_b200_v0.so
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r02_smi.log 2>&1
echo "== pytest" ; timeout 900 python -m pytest tests -m gpu -x -q > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r02_pytest_gpu.log
echo "== stress v1"
for i in 1 2; do timeout 300 python tools/stress.py wan_norm 4000 >> $O/r02_stress_v1.log 2>&1; echo "rc=$?" >> $O/r02_stress_v1.log; done
timeout 300 python tools/stress.py headline 4000 >> $O/r02_stress_v1.log 2>&1; echo "rc=$?" >> $O/r02_stress_v1.log
timeout 300 python tools/stress.py dit64 4000 >> $O/r02_stress_v1.log 2>&1; echo "rc=$?" >> $O/r02_stress_v1.log
timeout 300 python tools/stress.py wan 4000 >> $O/r02_stress_v1.log 2>&1; echo "rc=$?" >> $O/r02_stress_v1.log
tail -20 $O/r02_stress_v1.log
echo "== perf A/B"
for lib in v0 v1; do
  if [ $lib = v0 ]; then export MHLA_B200_LIB=$V0; else unset MHLA_B200_LIB; fi
  timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 > $O/r02_bench_$lib.json 2> $O/r02_bench_$lib.err
  timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 --no-normalize > $O/r02_bench_${lib}_nonorm.json 2>> $O/r02_bench_$lib.err
  timeout 120 python tools/prof_roles.py > $O/r02_prof_roles_$lib.log 2>&1
  timeout 120 python tools/timeline.py > $O/r02_timeline_$lib.log 2>&1
  timeout 120 python tools/trace_cta0.py > /dev/null 2>&1; mv $O/trace_cta0_norm.csv $O/r02_trace_cta0_$lib.csv
done
unset MHLA_B200_LIB
for knob in "MHLA_QKEEP=8" "MHLA_QKEEP=16" "MHLA_WSHINT=0" "MHLA_RUNAHEAD=3" "MHLA_SLOTS=1"; do
  echo "knob $knob" >> $O/r02_knobs.log
  env $knob timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 1 2>&1 | python -c "import sys,json; [print(json.loads(l)['ms_per_step'], json.loads(l)['roofline']['frac']) for l in sys.stdin if l.startswith('{')]" >> $O/r02_knobs.log 2>&1
done
cat $O/r02_knobs.log
grep -h ms_per_step $O/r02_bench_v*.json | python -c "import sys,json; [print(json.loads(l)['ms_per_step'], json.loads(l)['roofline']['frac']) for l in sys.stdin]"
echo "== configs"; timeout 600 python tools/bench_configs.py > $O/r02_configs_v1.jsonl 2>&1; cat $O/r02_configs_v1.jsonl
echo "== comparators"; timeout 900 python bench.py --impl reference-gpu --sweep > $O/r02_comparators_stdout.json 2> $O/r02_comparators.err; tail -3 $O/r02_comparators.err
echo "== sanitizer"
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_case.py --causal > $O/r02_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 $O/r02_sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool synccheck python tools/sanitize_case.py 0 > $O/r02_sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -5 $O/r02_sanitizer_synccheck.log
timeout 240 compute-sanitizer --tool racecheck python tools/sanitize_case.py 0 > $O/r02_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -8 $O/r02_sanitizer_racecheck.log
echo "== done"
