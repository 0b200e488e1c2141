#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_sharded_nccl.py -m gpu -q -rs -x > $O/r02_pytest_nccl_2gpu_b.log 2>&1; echo "nccl rc=$?"; tail -12 $O/r02_pytest_nccl_2gpu_b.log
