#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_modules.py -m gpu -q > $O/r02_pytest_wan.log 2>&1; echo "modules rc=$?"; tail -30 $O/r02_pytest_wan.log
timeout 900 python -m pytest tests -m gpu -q > $O/r02_pytest_gpu20.log 2>&1; echo "all rc=$?"; tail -4 $O/r02_pytest_gpu20.log
timeout 300 python tools/wan_layer_bench.py 2>&1 | tee $O/r02_wan_layer.log
