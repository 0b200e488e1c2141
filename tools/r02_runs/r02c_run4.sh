#!/bin/bash
# round 2, third session, run 4: which CTAs are slow?  per-CTA item counts / waits, then the item trace of the slowest and the fastest CTA
set -u
O=gpurun_out; mkdir -p $O
timeout 120 python tools/prof_roles.py --per-cta > $O/r02c_prof_per_cta.log 2>&1; tail -3 $O/r02c_prof_per_cta.log
read SLOW FAST < $O/slow_fast_cta.txt
MHLA_TRACE_CTA=$SLOW timeout 120 python tools/trace_cta0.py 2>&1 | tail -1
MHLA_TRACE_CTA=$FAST timeout 120 python tools/trace_cta0.py 2>&1 | tail -1
nvidia-smi --query-gpu=name,clocks.sm,clocks.mem,power.draw --format=csv
