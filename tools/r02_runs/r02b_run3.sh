#!/bin/bash
# round 2, second session, run 3: full -m gpu suite with the ABI v4 epilogue (gate / add), cfg timings, backward timings
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/r02b_pytest_gpu3.log 2>&1; echo "pytest rc=$?"; tail -25 $O/r02b_pytest_gpu3.log
timeout 600 python tools/bench_configs.py 2>/dev/null | tee $O/r02b_configs3.jsonl | cut -c1-150
timeout 300 python tools/bwd_bench.py > $O/r02b_bwd_bench3.jsonl 2> $O/r02b_bwd_bench3.err; timeout 300 python tools/bwd_bench.py --wan >> $O/r02b_bwd_bench3.jsonl 2>> $O/r02b_bwd_bench3.err; cut -c1-400 $O/r02b_bwd_bench3.jsonl; tail -3 $O/r02b_bwd_bench3.err
