#!/bin/bash
# GPU call 2: catch the rope+normaliser launch failure with a GPU core dump (cuda-gdb) and memcheck at full Wan size;
# A/B of kernel variants; the new parity tests.
set -u
O=gpurun_out; mkdir -p $O
export CUDA_ENABLE_COREDUMP_ON_EXCEPTION=1
export CUDA_COREDUMP_GENERATION_FLAGS='skip_global_memory,skip_local_memory,skip_constbank_memory'
for lib in v2 v1; do
  if [ $lib = v1 ]; then export MHLA_B200_LIB=$PWD/mhla_b200/libmhla_b200_v1.so; else unset MHLA_B200_LIB; fi
  rm -f /tmp/core_$lib*
  CUDA_COREDUMP_FILE=/tmp/core_${lib}_%p timeout 600 python tools/stress.py wan_norm 3000 > $O/r02_core_stress_$lib.log 2>&1
  echo "stress $lib rc=$?"; tail -3 $O/r02_core_stress_$lib.log
  for f in /tmp/core_${lib}_*; do
    [ -f "$f" ] || continue
    ls -la $f
    timeout 300 cuda-gdb -batch -ex "target cudacore $f" -ex "info cuda kernels" -ex "bt" -ex "info cuda lanes" -ex "x/24i \$pc-192" -ex "info registers system" > $O/r02_coredump_$lib.txt 2>&1
    head -60 $O/r02_coredump_$lib.txt
    break
  done
done
unset MHLA_B200_LIB CUDA_ENABLE_COREDUMP_ON_EXCEPTION
echo "== memcheck full size"
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/stress.py wan_norm 40 > $O/r02_memcheck_wan.log 2>&1; echo "memcheck rc=$?"; grep -v "^=========     " $O/r02_memcheck_wan.log | tail -25
echo "== perf"
for cfg in "v0:" "v1:" "v2:" "v2:MHLA_MIX_HI_ONLY=1" "v2:MHLA_RUNAHEAD=3"; do
  lib=${cfg%%:*}; knob=${cfg#*:}
  if [ $lib = v2 ]; then unset MHLA_B200_LIB; else export MHLA_B200_LIB=$PWD/mhla_b200/libmhla_b200_$lib.so; fi
  echo "$cfg" >> $O/r02_perf2.log
  env $knob timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 1 2>&1 | python -c "import sys,json; [print(json.loads(l)['ms_per_step'], json.loads(l)['roofline']['frac']) for l in sys.stdin if l.startswith('{')]" >> $O/r02_perf2.log 2>&1
done
unset MHLA_B200_LIB
cat $O/r02_perf2.log
timeout 120 python tools/prof_roles.py > $O/r02_prof_roles_v2.log 2>&1
timeout 120 python tools/timeline.py > $O/r02_timeline_v2.log 2>&1; cat $O/r02_timeline_v2.log
timeout 120 python tools/trace_cta0.py > /dev/null 2>&1; mv $O/trace_cta0_norm.csv $O/r02_trace_cta0_v2.csv
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q > $O/r02_pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -30 $O/r02_pytest_gpu2.log
echo "== done"
