#!/bin/bash
# session 3, run I: warp-per-point pair search as the default: whole GPU suite, bench, one parity sweep
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 200 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/tests_s3i.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt
tail -6 gpurun_out/tests_s3i.log
timeout 100 python bench.py --steps 3 --warmup 3 --cpu-chunks 0 --batch 128 > gpurun_out/bench_s3i.json 2> gpurun_out/bench_s3i.err; echo "bench exit $?" >> gpurun_out/summary.txt
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_s3i.json'))
    sm=d['config']['stage_ms_one_step']
    print('value %.1f'%d['value'],'ms %.2f'%d['ms_per_step'],'e2e %.1f'%d['e2e']['value'],'aff %.2f mv %.2f part %.2f'%(sm['affinity'],sm['matvec'],sm['partition']),'frac %.3f'%d['roofline']['frac'],'steps',d['config']['lanczos_steps_per_chunk'],'seg',d['config']['segments_per_chunk'],'unconv',d['config']['unconverged_nodes'])
except Exception as ex: print('failed',ex)
PY
timeout 60 python tools/parity_sweep.py --config tarl_spatial --chunks 32 --n-target 8192 --seed 7000 --oracle-cache parity_cache --out gpurun_out/parity_tarl_spatial.json > gpurun_out/parity_tarl_spatial.log 2>&1; echo "parity exit $?" >> gpurun_out/summary.txt
tail -1 gpurun_out/parity_tarl_spatial.log | cut -c1-250
cat gpurun_out/summary.txt
