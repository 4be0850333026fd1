#!/bin/bash
# session 2, run C: L2 bulk prefetch distance sweep (ANCUTS_X bits 4..7) on the bench workload
mkdir -p gpurun_out
ANCUTS_X=33 timeout 600 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/tests_x33.log 2>&1; echo "tests x33 exit $?" > gpurun_out/summary.txt
tail -3 gpurun_out/tests_x33.log
run() {
  ANCUTS_X=$1 timeout 600 python bench.py --steps 3 --warmup 3 --cpu-chunks 0 --batch $2 > gpurun_out/bench_x$1_b$2.json 2> gpurun_out/bench_x$1_b$2.err; echo "bench x$1 b$2 exit $?" >> gpurun_out/summary.txt
  python - $1 $2 <<'PY'
import json,sys
d=json.load(open('gpurun_out/bench_x%s_b%s.json'%(sys.argv[1],sys.argv[2])))
print('X',sys.argv[1],'batch',sys.argv[2],'value %.1f'%d['value'],'ms %.2f'%d['ms_per_step'],'e2e %.1f'%d['e2e']['value'],'matvec_ms %.2f'%d['config']['stage_ms_one_step']['matvec'],'frac %.3f'%d['roofline']['frac'])
PY
}
for x in 1 33 65 129; do run $x 64; done
for x in 1 33 65; do run $x 16; done
run 65 128
cat gpurun_out/summary.txt
