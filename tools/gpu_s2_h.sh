#!/bin/bash
# session 2, run H: TMA ring v2 (4 KB stages, basis mostly in L2) vs plain; basis-in-global cost
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
ANCUTS_X=9217 timeout 600 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/tests_x9217.log 2>&1; echo "tests x9217 exit $?" >> gpurun_out/summary.txt
tail -5 gpurun_out/tests_x9217.log
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
run 1065 64    # reference point
run 5161 64    # same, basis rows in global memory only
run 9217 64    # TMA ring v2 + GS1
run 9217 16
run 9217 128
ANCUTS_X=9217 ANCUTS_PHASES=1 timeout 300 python tools/level_profile.py --batch 64 --out gpurun_out/levels_x9217_b64.json > gpurun_out/levels_x9217_b64.log 2>&1
grep "cluster size" gpurun_out/levels_x9217_b64.log
cat gpurun_out/summary.txt
