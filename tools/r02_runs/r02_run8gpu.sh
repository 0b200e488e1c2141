#!/bin/bash
# 8-GPU call: NCCL sharding tests at world 2/4/8 (bitwise vs single GPU) and the strong-scaling bench at N=8 and N=4
set -u
O=gpurun_out; mkdir -p $O
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_sharded_nccl.py -m gpu -q -rs > $O/r02_pytest_nccl_8gpu.log 2>&1; echo "nccl rc=$?"; tail -6 $O/r02_pytest_nccl_8gpu.log
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > $O/r02_bench_${n}gpu.json 2> $O/r02_bench_${n}gpu.err; echo "bench$n rc=$?"; python -c "
import json,sys
l=[x for x in open('$O/r02_bench_${n}gpu.json') if x.startswith('{')][-1]; r=json.loads(l)
print($n, 'value', r['value'], 'ms', r['ms_per_step'], 'gather', r.get('with_gather'), 'e2e', r['e2e']['value'], r['config']['launch'])"
done
