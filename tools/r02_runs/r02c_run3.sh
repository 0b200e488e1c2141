#!/bin/bash
# round 2, third session, run 3: coalesced gate/add epilogue (parity + operator timing), role profile and CTA trace of the new default
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_blockmix_gpu.py tests/test_modules.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python tools/wan_layer_bench.py 2>&1 | tail -3
timeout 120 python tools/prof_roles.py > $O/r02c_prof_roles.log 2>&1; cat $O/r02c_prof_roles.log
timeout 120 python tools/trace_cta0.py > $O/r02c_trace.log 2>&1; tail -2 $O/r02c_trace.log
timeout 120 python tools/timeline.py > $O/r02c_timeline.log 2>&1; cat $O/r02c_timeline.log
