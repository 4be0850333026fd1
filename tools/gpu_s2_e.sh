#!/bin/bash
# session 2, run E: TMA ring matvec (ANCUTS_X=9 / 11) vs L2 prefetch (33)
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
ANCUTS_X=9 timeout 600 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/tests_x9.log 2>&1; echo "tests x9 exit $?" >> gpurun_out/summary.txt
tail -3 gpurun_out/tests_x9.log
run() {
  ANCUTS_X=$1 timeout 600 python bench.py --steps 3 --warmup 3 --cpu-chunks 0 --batch $2 > gpurun_out/bench_x$1_b$2.json 2> gpurun_out/bench_x$1_b$2.err; echo "bench x$1 b$2 exit $?" >> gpurun_out/summary.txt
  python - $1 $2 <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/bench_x%s_b%s.json'%(sys.argv[1],sys.argv[2])))
    print('X',sys.argv[1],'batch',sys.argv[2],'value %.1f'%d['value'],'ms %.2f'%d['ms_per_step'],'e2e %.1f'%d['e2e']['value'],'matvec_ms %.2f'%d['config']['stage_ms_one_step']['matvec'],'frac %.3f'%d['roofline']['frac'],'unconv',d['config']['unconverged_nodes'],'steps',d['config']['lanczos_steps_per_chunk'])
except Exception as ex: print('failed',sys.argv[1:],ex)
PY
}
run 9 64; run 11 64; run 33 64; run 9 16; run 33 16
ANCUTS_X=9 ANCUTS_PHASES=1 timeout 300 python tools/level_profile.py --batch 64 --out gpurun_out/levels_x9_b64.json > gpurun_out/levels_x9_b64.log 2>&1; echo "levels exit $?" >> gpurun_out/summary.txt
grep "cluster size" gpurun_out/levels_x9_b64.log
cat gpurun_out/summary.txt
