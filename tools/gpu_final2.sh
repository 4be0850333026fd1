#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "tests exit $?" > gpurun_out/summary.txt
tail -4 gpurun_out/tests.log
CMD="python bench.py --steps 1 --warmup 1 --cpu-chunks 0"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_lanczos_cluster -s 0 -c 6 -o gpurun_out/prof_cluster_final $CMD > gpurun_out/ncu_cluster.log 2>&1
echo "cluster capture exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
