#!/bin/bash
# session 2, run L: wide-tile affinity pass 1 with tile-local union-find; store ceiling; default flags (ring + mix)
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt
tail -3 gpurun_out/tests.log

timeout 300 python tools/tc_bench.py > gpurun_out/tc_bench.log 2>&1; echo "tc_bench exit $?" >> gpurun_out/summary.txt
grep '"impl": 0' gpurun_out/tc_bench.log
run() {
  timeout 600 python bench.py --steps 3 --warmup 3 --cpu-chunks 0 --batch $1 > gpurun_out/bench_b$1.json 2> gpurun_out/bench_b$1.err; echo "bench b$1 exit $?" >> gpurun_out/summary.txt
  python - $1 <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/bench_b%s.json'%(sys.argv[1])))
    sm=d['config']['stage_ms_one_step']
    print('batch',sys.argv[1],'value %.1f'%d['value'],'ms %.2f'%d['ms_per_step'],'e2e %.1f'%d['e2e']['value'],'aff %.2f mv %.2f part %.2f'%(sm['affinity'],sm['matvec'],sm['partition']),'frac %.3f'%d['roofline']['frac'],'steps',d['config']['lanczos_steps_per_chunk'],'seg',d['config']['segments_per_chunk'],'unconv',d['config']['unconverged_nodes'])
except Exception as ex: print('failed',sys.argv[1:],ex)
PY
}
run 64; run 128
cat gpurun_out/summary.txt
