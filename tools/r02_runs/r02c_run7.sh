#!/bin/bash
# round 2, third session, run 7: device time of wan_prep (CUDA graph, host cost off the path) for the staged-ring kernel
set -u
O=gpurun_out; mkdir -p $O
timeout 300 python tools/wan_layer_bench.py > $O/r02c_wan_layer3.log 2>&1; cat $O/r02c_wan_layer3.log | tail -6
