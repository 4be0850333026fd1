#!/bin/bash
# session 2, run F: software-pipelined matvec (ANCUTS_X 9 / 41) and two-pass affinity (default on; bit 8 = off)
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
ANCUTS_X=41 timeout 600 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/tests_x41.log 2>&1; echo "tests x41 exit $?" >> gpurun_out/summary.txt
tail -5 gpurun_out/tests_x41.log
run() {
  ANCUTS_X=$1 timeout 600 python bench.py --steps 3 --warmup 3 --cpu-chunks 0 --batch $2 > gpurun_out/bench_x$1_b$2.json 2> gpurun_out/bench_x$1_b$2.err; echo "bench x$1 b$2 exit $?" >> gpurun_out/summary.txt
  python - $1 $2 <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/bench_x%s_b%s.json'%(sys.argv[1],sys.argv[2])))
    sm=d['config']['stage_ms_one_step']
    print('X',sys.argv[1],'batch',sys.argv[2],'value %.1f'%d['value'],'ms %.2f'%d['ms_per_step'],'e2e %.1f'%d['e2e']['value'],'aff %.2f mv %.2f part %.2f'%(sm['affinity'],sm['matvec'],sm['partition']),'frac %.3f'%d['roofline']['frac'],'steps',d['config']['lanczos_steps_per_chunk'],'seg',d['config']['segments_per_chunk'])
except Exception as ex: print('failed',sys.argv[1:],ex)
PY
}
run 289 64   # 256+33: one-kernel affinity, plain matvec + prefetch 2
run 33 64    # two-pass affinity, plain matvec + prefetch 2
run 9 64     # pipelined matvec, no prefetch
run 41 64    # pipelined matvec + next-step prefetch
run 41 16
run 33 16
timeout 300 python tools/tc_bench.py > gpurun_out/tc_bench.log 2>&1; echo "tc_bench exit $?" >> gpurun_out/summary.txt
grep '"impl": 0' gpurun_out/tc_bench.log
cat gpurun_out/summary.txt
