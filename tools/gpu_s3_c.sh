#!/bin/bash
# session 3, run C: pipelined Sturm counts, 128 vs 256 shifts per round (ANCUTS_X bit 16), feature pooling row N3
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/tests_s3c.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt
tail -15 gpurun_out/tests_s3c.log
run() {
  ANCUTS_X=$1 timeout 600 python bench.py --steps 3 --warmup 3 --cpu-chunks 0 --batch $2 > gpurun_out/bench_x$1_b$2.json 2> gpurun_out/bench_x$1_b$2.err; echo "bench x$1 b$2 exit $?" >> gpurun_out/summary.txt
  python - $1 $2 <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/bench_x%s_b%s.json'%(sys.argv[1],sys.argv[2])))
    sm=d['config']['stage_ms_one_step']
    print('x',sys.argv[1],'batch',sys.argv[2],'value %.1f'%d['value'],'ms %.2f'%d['ms_per_step'],'e2e %.1f'%d['e2e']['value'],'aff %.2f mv %.2f part %.2f'%(sm['affinity'],sm['matvec'],sm['partition']),'frac %.3f'%d['roofline']['frac'],'steps',d['config']['lanczos_steps_per_chunk'],'seg',d['config']['segments_per_chunk'],'unconv',d['config']['unconverged_nodes'])
except Exception as ex: print('failed',sys.argv[1:],ex)
PY
}
run 58370 128; run 123906 128
for x in 58370 123906; do
  ANCUTS_X=$x ANCUTS_PHASES=1 timeout 400 python tools/level_profile.py --batch 128 --out gpurun_out/levels_x$x.json > gpurun_out/levels_x$x.log 2>&1; echo "levels x$x exit $?" >> gpurun_out/summary.txt
  grep "cluster size" gpurun_out/levels_x$x.log
done
timeout 600 python tools/pool_bench.py --out gpurun_out/pool_bench.json > gpurun_out/pool_bench.log 2>&1; echo "pool bench exit $?" >> gpurun_out/summary.txt
tail -2 gpurun_out/pool_bench.log
cat gpurun_out/summary.txt
