#!/bin/bash
# round 2, second session, run 4: gated-norm kernel, block_wsum in the backward, Wan layer with fused gate / lepe, training step
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_backward_gpu.py tests/test_causal_gpu.py tests/test_modules.py -m gpu -q -x > $O/r02b_pytest_gpu4.log 2>&1; echo "pytest rc=$?"; tail -12 $O/r02b_pytest_gpu4.log
timeout 300 python tools/bwd_bench.py > $O/r02b_bwd_bench4.jsonl 2> $O/r02b_bwd_bench4.err; timeout 300 python tools/bwd_bench.py --wan >> $O/r02b_bwd_bench4.jsonl 2>> $O/r02b_bwd_bench4.err; cut -c1-330 $O/r02b_bwd_bench4.jsonl; tail -3 $O/r02b_bwd_bench4.err
timeout 600 python tools/wan_layer_bench.py > $O/r02b_wan_layer4.log 2>&1; tail -6 $O/r02b_wan_layer4.log
