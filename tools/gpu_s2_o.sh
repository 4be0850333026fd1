#!/bin/bash
# session 2, run O: ncu full capture of the two affinity passes (stage entry point, N = 8.6 k)
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_affinity -s 3 -c 1 -o /tmp/prof_aff python tools/aff_once.py > gpurun_out/ncu_aff.log 2>&1
echo "aff capture exit $?" >> gpurun_out/summary.txt
ncu -i /tmp/prof_aff.ncu-rep --page raw --csv > gpurun_out/ncu_aff_raw.csv 2>/dev/null
ncu -i /tmp/prof_aff.ncu-rep --page source --csv > gpurun_out/ncu_aff_source.csv 2>/dev/null
ls -la gpurun_out/ncu_aff_*.csv
cat gpurun_out/summary.txt
