#!/bin/bash
# round 2, third session, run 6: wan_prep with two rows of raw look-ahead per thread - parity (module tests) + timing
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_modules.py tests/test_blockmix_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python tools/wan_layer_bench.py > $O/r02c_wan_layer2.log 2>&1; cat $O/r02c_wan_layer2.log | tail -7
