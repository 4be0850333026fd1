#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "tests exit $?" > gpurun_out/summary.txt
tail -3 gpurun_out/tests.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 900 python tools/parity_sweep.py --config tarl_spatial --chunks 48 --n-target 4096 --out gpurun_out/parity_tarl.json > gpurun_out/parity_tarl.log 2>&1; echo "parity exit $?" >> gpurun_out/summary.txt
tail -2 gpurun_out/parity_tarl.log
cat gpurun_out/summary.txt
