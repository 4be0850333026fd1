#!/bin/bash
# session 3 evidence run: tests, parity sweeps (oracle labels cached on the CPU box), map-level metrics, ncu launch list and DRAM
# bytes of the cluster kernels (copied to profiles/ BEFORE the bench so that roofline.traffic matches this build), one full
# capture, bench (both arms), pooling row
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt
tail -3 gpurun_out/tests.log
for cfg in spatial tarl_spatial tarl_spatial_dino; do
  timeout 900 python tools/parity_sweep.py --config $cfg --chunks 32 --n-target 8192 --seed 7000 --oracle-cache parity_cache --out gpurun_out/parity_$cfg.json > gpurun_out/parity_$cfg.log 2>&1; echo "parity $cfg exit $?" >> gpurun_out/summary.txt
  tail -1 gpurun_out/parity_$cfg.log | cut -c1-300
done
timeout 900 python tools/parity_sweep.py --config tarl_spatial --chunks 6 --n-target 16384 --seed 7200 --oracle-cache parity_cache --out gpurun_out/parity_tarl_spatial_16k.json > gpurun_out/parity_16k.log 2>&1; echo "parity 16k exit $?" >> gpurun_out/summary.txt
tail -1 gpurun_out/parity_16k.log | cut -c1-300
timeout 900 python tools/parity_sweep.py --config tarl_spatial --chunks 16 --n-target 8192 --seed 7400 --clutter 40 --oracle-cache parity_cache --out gpurun_out/parity_tarl_spatial_clutter40.json > gpurun_out/parity_clutter.log 2>&1; echo "parity clutter exit $?" >> gpurun_out/summary.txt
tail -1 gpurun_out/parity_clutter.log | cut -c1-400
timeout 900 python tools/map_eval.py --chunks 24 --n-per-chunk 6000 --out gpurun_out/map_eval.json > gpurun_out/map_eval.log 2>&1; echo "map exit $?" >> gpurun_out/summary.txt
tail -1 gpurun_out/map_eval.log | cut -c1-600
CMD="python tools/one_step.py --batch 128"
$CMD > gpurun_out/one_step.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?" >> gpurun_out/summary.txt
python tools/ncu_summarise.py gpurun_out/launches.csv gpurun_out/launch_list.csv | head -14
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_lanczos_cluster -c 200 --csv --log-file gpurun_out/cluster_dram.csv $CMD > gpurun_out/ncu_dram.log 2>&1
echo "dram list exit $?" >> gpurun_out/summary.txt
python tools/ncu_summarise.py gpurun_out/cluster_dram.csv gpurun_out/cluster_dram_summary.csv && cp gpurun_out/cluster_dram_summary.csv profiles/r1c_cluster_dram_b128.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lanczos_cluster -s 1 -c 2 -o /tmp/prof_cluster python tools/one_step.py --batch 128 --passes 1 > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?" >> gpurun_out/summary.txt
ncu -i /tmp/prof_cluster.ncu-rep --page raw --csv > gpurun_out/ncu_cluster_raw.csv 2>/dev/null
ls -la /tmp/prof_cluster.ncu-rep && cp /tmp/prof_cluster.ncu-rep gpurun_out/prof_cluster_s3.ncu-rep
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
head -c 700 gpurun_out/bench.json; echo
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm exit $?" >> gpurun_out/summary.txt
head -c 400 gpurun_out/bench_reference.json; echo
timeout 600 python tools/pool_bench.py --out gpurun_out/pool_bench.json > gpurun_out/pool_bench.log 2>&1; echo "pool bench exit $?" >> gpurun_out/summary.txt
tail -1 gpurun_out/pool_bench.log
cat gpurun_out/summary.txt
