#!/bin/bash
# session 2, run S: per-bin gather launches; union-find placement A/B (bit 9 = one global union per pair)
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt
tail -3 gpurun_out/tests.log
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
run 9219 128; run 9731 128; run 9219 64; run 9731 64
cat gpurun_out/summary.txt
