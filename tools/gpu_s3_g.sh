#!/bin/bash
# session 3, run G: pairs of the deferred affinity from a cell grid (ANCUTS_X bit 19) instead of the tile sweep
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/tests_s3g.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt
tail -12 gpurun_out/tests_s3g.log
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
run 517122 128; run 1041410 128
for cfg in tarl_spatial tarl_spatial_dino; do
  timeout 900 python tools/parity_sweep.py --config $cfg --chunks 32 --n-target 8192 --seed 7000 --oracle-cache parity_cache --out gpurun_out/parity_$cfg.json > gpurun_out/parity_$cfg.log 2>&1; echo "parity $cfg exit $?" >> gpurun_out/summary.txt
  tail -1 gpurun_out/parity_$cfg.log | cut -c1-330
done
timeout 900 python tools/parity_sweep.py --config tarl_spatial --chunks 6 --n-target 16384 --seed 7200 --oracle-cache parity_cache --out gpurun_out/parity_tarl_spatial_16k.json > gpurun_out/parity_16k.log 2>&1; echo "parity 16k exit $?" >> gpurun_out/summary.txt
tail -1 gpurun_out/parity_16k.log | cut -c1-300
timeout 900 python tools/parity_sweep.py --config tarl_spatial --chunks 16 --n-target 8192 --seed 7400 --clutter 40 --oracle-cache parity_cache --out gpurun_out/parity_tarl_spatial_clutter40.json > gpurun_out/parity_clutter.log 2>&1; echo "parity clutter exit $?" >> gpurun_out/summary.txt
tail -1 gpurun_out/parity_clutter.log | cut -c1-330
cat gpurun_out/summary.txt
