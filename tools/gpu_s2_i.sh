#!/bin/bash
# session 2, run I: ncu full capture of the cluster kernels (level 0, batch 64), exported to CSV on the box (the reports exceed the 64 MiB return limit)
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
CMD="python bench.py --steps 1 --warmup 1 --cpu-chunks 0 --batch 64"
cap() {  # name, ANCUTS_X
  ANCUTS_X=$2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_lanczos_cluster -s 1 -c 2 -o /tmp/prof_$1 $CMD > gpurun_out/ncu_$1.log 2>&1
  echo "$1 capture exit $?" >> gpurun_out/summary.txt
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/ncu_$1_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page source --csv > gpurun_out/ncu_$1_source.csv 2>/dev/null
  ls -la /tmp/prof_$1.ncu-rep gpurun_out/ncu_$1_*.csv
}
cap ring 9217
cap plain 1065
cat gpurun_out/summary.txt
