#!/bin/bash
# launch list + one full capture of the dominant kernel, each only after the plain run exited 0
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --batch 8 --cpu-chunks 0"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 9000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_matvec -s 300 -c 3 -o gpurun_out/prof_matvec $CMD > gpurun_out/ncu_matvec.log 2>&1
echo "matvec capture exit $?"
ls -la gpurun_out
