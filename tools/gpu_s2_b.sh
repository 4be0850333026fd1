#!/bin/bash
# session 2, run B: per-level trace, matvec variants (ANCUTS_X), ncu full capture of the cluster kernels at batch 64
mkdir -p gpurun_out
ANCUTS_X=3 timeout 600 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/tests_x3.log 2>&1; echo "tests x3 exit $?" > gpurun_out/summary.txt
tail -3 gpurun_out/tests_x3.log
timeout 300 python tools/level_profile.py --batch 16 --out gpurun_out/levels_b16.json > gpurun_out/levels_b16.log 2>&1; echo "levels16 exit $?" >> gpurun_out/summary.txt
timeout 300 python tools/level_profile.py --batch 64 --out gpurun_out/levels_b64.json > gpurun_out/levels_b64.log 2>&1; echo "levels64 exit $?" >> gpurun_out/summary.txt
tail -2 gpurun_out/levels_b64.log
for x in 0 1 3; do
  ANCUTS_X=$x timeout 600 python bench.py --steps 3 --warmup 3 --cpu-chunks 0 --batch 64 > gpurun_out/bench_x$x.json 2> gpurun_out/bench_x$x.err; echo "bench x$x exit $?" >> gpurun_out/summary.txt
  python - $x <<'PY'
import json,sys
d=json.load(open('gpurun_out/bench_x%s.json'%sys.argv[1]))
print('X',sys.argv[1],'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
print('stage_ms',d['config']['stage_ms_one_step'])
PY
done
CMD="python bench.py --steps 1 --warmup 1 --cpu-chunks 0 --batch 64"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_lanczos_cluster -s 0 -c 5 -o gpurun_out/prof_cluster_b64 $CMD > gpurun_out/ncu_cluster.log 2>&1
echo "cluster capture exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
