#!/bin/bash
# session 3, run H: bench and ncu launch list of the default build (deferred affinity, tile pairs); kernel times of the grid pair search
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
head -c 600 gpurun_out/bench.json; echo
CMD="python tools/one_step.py --batch 128"
$CMD > gpurun_out/one_step.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?" >> gpurun_out/summary.txt
python tools/ncu_summarise.py gpurun_out/launches.csv gpurun_out/launch_list.csv | head -16
ANCUTS_X=1041410 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_pair -c 40 --csv --log-file gpurun_out/pair_kernels.csv python tools/one_step.py --batch 8 > gpurun_out/ncu_pair.log 2>&1
echo "pair kernels exit $?" >> gpurun_out/summary.txt
python tools/ncu_summarise.py gpurun_out/pair_kernels.csv gpurun_out/pair_kernels_summary.csv
cat gpurun_out/summary.txt
