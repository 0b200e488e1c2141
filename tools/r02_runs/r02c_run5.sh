#!/bin/bash
# round 2, third session, run 5: four accumulator buffers (MHLA_ACC4), backwards readout with a tail (MHLA_REVERSE3), static first
# item for D = 128 (new default) - A/B, then the full -m gpu suite + bench on the default build AND on the ACC4 build
set -u
O=gpurun_out; mkdir -p $O
L=$PWD/mhla_b200
timeout 600 python tools/ab_libs.py --reps 3 --shapes headline,nonorm,wan,n8k base=$L/libmhla_b200.so acc4=$L/libmhla_b200_acc4.so rev3=$L/libmhla_b200_rev3.so@MHLA_REVERSE3=3 rev5=$L/libmhla_b200_rev5.so@MHLA_REVERSE3=5 acc4rev=$L/libmhla_b200_acc4rev.so@MHLA_REVERSE3=3 old=$L/libmhla_b200_old.so > $O/r02c_ab_3.log 2>&1
cut -c1-160 $O/r02c_ab_3.log
show() { python -c "
import json,sys
r=json.loads([x for x in open('$1') if x.startswith('{')][-1])
print('$1', 'ms', round(r['ms_per_step'],5), 'frac', round(r['roofline']['frac'],4), 'e2e ms', round(r['e2e']['ms_per_step'],3), r['clocks'])"; }
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -q > $O/r02c_pytest_gpu_final2.log 2>&1; echo "pytest(base) rc=$?"; tail -3 $O/r02c_pytest_gpu_final2.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02c_bench_final2.json 2> $O/r02c_bench_final2.err; show $O/r02c_bench_final2.json
MHLA_B200_LIB=$L/libmhla_b200_acc4.so timeout 900 python -m pytest tests -m gpu -q > $O/r02c_pytest_gpu_acc4.log 2>&1; echo "pytest(acc4) rc=$?"; tail -3 $O/r02c_pytest_gpu_acc4.log
MHLA_B200_LIB=$L/libmhla_b200_acc4.so timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02c_bench_acc4.json 2> $O/r02c_bench_acc4.err; show $O/r02c_bench_acc4.json
for v in "headline 4000" "rn_d64 3000" "dit64 3000"; do MHLA_B200_LIB=$L/libmhla_b200_acc4.so MHLA_STRESS_PRODUCT=1 timeout 300 python tools/stress.py $v 2>&1 | tail -1; done
