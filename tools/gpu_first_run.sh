#!/bin/bash
# first GPU bring-up: every test file in its own process so one CUDA fault does not hide the rest
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for f in tests/test_gpu_stages.py tests/test_gpu_segment.py; do
  b=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/$b.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt
  tail -5 gpurun_out/$b.log
done
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
