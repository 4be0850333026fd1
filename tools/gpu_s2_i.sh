#!/bin/bash
# session 2, run I: ncu full capture of the ring-v2 cluster kernels (level 0, batch 64)
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
CMD="python bench.py --steps 1 --warmup 1 --cpu-chunks 0 --batch 64"
ANCUTS_X=9217 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_lanczos_cluster -s 0 -c 6 -o gpurun_out/prof_ring_b64 $CMD > gpurun_out/ncu_ring.log 2>&1
echo "ring capture exit $?" >> gpurun_out/summary.txt
ANCUTS_X=1065 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_lanczos_cluster -s 0 -c 6 -o gpurun_out/prof_plain_b64 $CMD > gpurun_out/ncu_plain.log 2>&1
echo "plain capture exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
