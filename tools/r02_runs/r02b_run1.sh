#!/bin/bash
# round 2, second session, run 1: native backward (forward kernels with permuted operands) - parity + timing
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_backward_gpu.py -q -x > $O/r02b_pytest_backward.log 2>&1; echo "pytest rc=$?"; tail -15 $O/r02b_pytest_backward.log
timeout 300 python tools/bwd_bench.py > $O/r02b_bwd_bench.jsonl 2> $O/r02b_bwd_bench.err; cat $O/r02b_bwd_bench.jsonl; tail -3 $O/r02b_bwd_bench.err
