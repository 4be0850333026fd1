#!/bin/bash
# session 2, run G: single-pass Gram-Schmidt after the three-term recurrence (1024), per-lane prefetch (2048), FMA-chain matvec (8)
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
ANCUTS_X=1065 timeout 600 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/tests_x1065.log 2>&1; echo "tests x1065 exit $?" >> gpurun_out/summary.txt
tail -5 gpurun_out/tests_x1065.log
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
run 33 64      # baseline: plain + bulk prefetch 2
run 1057 64    # + single-pass GS
run 2081 64    # per-lane prefetch 2
run 41 64      # FMA chain + bulk prefetch 2
run 1065 64    # FMA chain + GS1 + bulk prefetch
run 1065 16
ANCUTS_X=1065 ANCUTS_PHASES=1 timeout 300 python tools/level_profile.py --batch 64 --out gpurun_out/levels_x1065_b64.json > gpurun_out/levels_x1065_b64.log 2>&1
grep "cluster size" gpurun_out/levels_x1065_b64.log
cat gpurun_out/summary.txt
