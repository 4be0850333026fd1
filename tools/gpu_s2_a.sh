#!/bin/bash
# session 2, run A: regression check + batch-size and lane scaling of the throughput
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "tests exit $?" > gpurun_out/summary.txt
tail -3 gpurun_out/tests.log
for b in 16 32 64; do
  timeout 600 python bench.py --steps 3 --warmup 3 --cpu-chunks 0 --batch $b > gpurun_out/bench_b$b.json 2> gpurun_out/bench_b$b.err; echo "bench b$b exit $?" >> gpurun_out/summary.txt
  python - $b <<'PY'
import json,sys
d=json.load(open('gpurun_out/bench_b%s.json'%sys.argv[1]))
print('batch',sys.argv[1],'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'])
print('stage_ms',d['config']['stage_ms_one_step'])
print('roofline',d['roofline']['achieved'],d['roofline']['frac'])
PY
done
timeout 600 python tools/lanes_bench.py > gpurun_out/lanes.log 2>&1; echo "lanes exit $?" >> gpurun_out/summary.txt
cat gpurun_out/lanes.log | tail -5
cat gpurun_out/summary.txt
