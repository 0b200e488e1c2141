#!/bin/bash
# 2-GPU call: NCCL sharding tests (bitwise vs single GPU) and the strong-scaling bench at N=2
set -u
O=gpurun_out; mkdir -p $O
nvidia-smi -L
timeout 900 python -m pytest tests/test_sharded_nccl.py -m gpu -q -rs > $O/r02_pytest_nccl_2gpu.log 2>&1; echo "nccl rc=$?"; tail -8 $O/r02_pytest_nccl_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r02_bench_2gpu.json 2> $O/r02_bench_2gpu.err; echo "bench2 rc=$?"; cat $O/r02_bench_2gpu.json; tail -3 $O/r02_bench_2gpu.err
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02_bench_1gpu.json 2> $O/r02_bench_1gpu.err; cat $O/r02_bench_1gpu.json
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r02_bench_reference.json 2>&1; cat $O/r02_bench_reference.json
