#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --batch 16 --cpu-chunks 0"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_lanczos_cluster -s 0 -c 6 -o gpurun_out/prof_cluster $CMD > gpurun_out/ncu_cluster.log 2>&1
echo "cluster capture exit $?"
tail -3 gpurun_out/ncu_cluster.log
