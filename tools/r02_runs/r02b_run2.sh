#!/bin/bash
# round 2, second session, run 2: native backward with the aux kernels, grid autograd, Wan training path
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_backward_gpu.py "tests/test_modules.py::test_wan_module_training_step_through_the_3d_block_view" tests/test_modules.py -q -x -m gpu > $O/r02b_pytest_backward2.log 2>&1; echo "pytest rc=$?"; tail -15 $O/r02b_pytest_backward2.log
timeout 300 python tools/bwd_bench.py > $O/r02b_bwd_bench2.jsonl 2> $O/r02b_bwd_bench2.err; cat $O/r02b_bwd_bench2.jsonl; tail -3 $O/r02b_bwd_bench2.err
timeout 300 python tools/bwd_bench.py --wan >> $O/r02b_bwd_bench2.jsonl 2>> $O/r02b_bwd_bench2.err; tail -1 $O/r02b_bwd_bench2.jsonl; tail -3 $O/r02b_bwd_bench2.err
