#!/bin/bash
# session 2, run J: mixed XU / integer float->double widening (bit 1) on the plain and the ring matvec
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
ANCUTS_X=9219 timeout 600 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/tests_x9219.log 2>&1; echo "tests x9219 exit $?" >> gpurun_out/summary.txt
tail -3 gpurun_out/tests_x9219.log
ANCUTS_X=1067 timeout 600 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/tests_x1067.log 2>&1; echo "tests x1067 exit $?" >> gpurun_out/summary.txt
tail -3 gpurun_out/tests_x1067.log
run() {
  ANCUTS_X=$1 timeout 600 python bench.py --steps 3 --warmup 3 --cpu-chunks 0 --batch $2 > gpurun_out/bench_x$1_b$2.json 2> gpurun_out/bench_x$1_b$2.err; echo "bench x$1 b$2 exit $?" >> gpurun_out/summary.txt
  python - $1 $2 <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/bench_x%s_b%s.json'%(sys.argv[1],sys.argv[2])))
    sm=d['config']['stage_ms_one_step']
    print('X',sys.argv[1],'batch',sys.argv[2],'value %.1f'%d['value'],'ms %.2f'%d['ms_per_step'],'e2e %.1f'%d['e2e']['value'],'aff %.2f mv %.2f part %.2f'%(sm['affinity'],sm['matvec'],sm['partition']),'frac %.3f'%d['roofline']['frac'],'steps',d['config']['lanczos_steps_per_chunk'],'seg',d['config']['segments_per_chunk'],'unconv',d['config']['unconverged_nodes'])
except Exception as ex: print('failed',sys.argv[1:],ex)
PY
}
run 1065 64; run 1067 64; run 9217 64; run 9219 64; run 9219 128; run 1067 128; run 9219 16; run 1067 16
cat gpurun_out/summary.txt
