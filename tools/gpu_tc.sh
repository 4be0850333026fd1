#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider -k "tensor_core" > gpurun_out/tests_tc.log 2>&1; echo "tc tests exit $?"
tail -5 gpurun_out/tests_tc.log
timeout 300 python tools/tc_bench.py > gpurun_out/tc_bench.log 2>&1; echo "tc bench exit $?"; cat gpurun_out/tc_bench.log | tail -8
