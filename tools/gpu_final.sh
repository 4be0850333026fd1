#!/bin/bash
# round-end evidence: tests, parity sweeps, map-level metrics, bench (both arms), ncu launch list + full capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "tests exit $?" > gpurun_out/summary.txt
tail -2 gpurun_out/tests.log
for cfg in spatial tarl_spatial tarl_spatial_dino; do
  timeout 1200 python tools/parity_sweep.py --config $cfg --chunks 32 --n-target 8192 --seed 7000 --out gpurun_out/parity_$cfg.json > gpurun_out/parity_$cfg.log 2>&1; echo "parity $cfg exit $?" >> gpurun_out/summary.txt
  tail -1 gpurun_out/parity_$cfg.log | cut -c1-420
done
timeout 1200 python tools/map_eval.py --chunks 24 --n-per-chunk 6000 --out gpurun_out/map_eval.json > gpurun_out/map_eval.log 2>&1; echo "map exit $?" >> gpurun_out/summary.txt
tail -1 gpurun_out/map_eval.log | cut -c1-900
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
head -c 400 gpurun_out/bench.json; echo
CMD="python bench.py --steps 1 --warmup 1 --cpu-chunks 0"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
