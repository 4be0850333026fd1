#!/bin/bash
# session 2, run D: phase clocks of the cluster kernel, prefetch distance 1/3, lanes at 64 chunks
mkdir -p gpurun_out
for b in 16 64; do
ANCUTS_X=33 ANCUTS_PHASES=1 timeout 300 python tools/level_profile.py --batch $b --out gpurun_out/levels_x33_b$b.json > gpurun_out/levels_x33_b$b.log 2>&1; echo "levels b$b exit $?" >> gpurun_out/summary.txt
grep "cluster size" gpurun_out/levels_x33_b$b.log
done
run() {
  ANCUTS_X=$1 timeout 600 python bench.py --steps 3 --warmup 3 --cpu-chunks 0 --batch $2 > gpurun_out/bench_x$1_b$2.json 2> gpurun_out/bench_x$1_b$2.err; echo "bench x$1 b$2 exit $?" >> gpurun_out/summary.txt
  python - $1 $2 <<'PY'
import json,sys
d=json.load(open('gpurun_out/bench_x%s_b%s.json'%(sys.argv[1],sys.argv[2])))
print('X',sys.argv[1],'batch',sys.argv[2],'value %.1f'%d['value'],'ms %.2f'%d['ms_per_step'],'e2e %.1f'%d['e2e']['value'],'matvec_ms %.2f'%d['config']['stage_ms_one_step']['matvec'],'frac %.3f'%d['roofline']['frac'])
PY
}
run 17 64; run 49 64; run 33 64
ANCUTS_X=33 timeout 600 python tools/lanes_bench.py --chunks 64 > gpurun_out/lanes64.log 2>&1; echo "lanes exit $?" >> gpurun_out/summary.txt
cat gpurun_out/lanes64.log | tail -4
cat gpurun_out/summary.txt
