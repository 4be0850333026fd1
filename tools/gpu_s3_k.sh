#!/bin/bash
# session 3, run K (last GPU seconds of the round): grid pair search with batched unions
mkdir -p gpurun_out
ANCUTS_X=1041410 timeout 80 python bench.py --steps 3 --warmup 3 --cpu-chunks 0 --batch 128 > gpurun_out/bench_s3k.json 2> gpurun_out/bench_s3k.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_s3k.json'))
    sm=d['config']['stage_ms_one_step']
    print('value %.1f'%d['value'],'ms %.2f'%d['ms_per_step'],'e2e %.1f'%d['e2e']['value'],'aff %.2f mv %.2f part %.2f'%(sm['affinity'],sm['matvec'],sm['partition']),'seg',d['config']['segments_per_chunk'],'unconv',d['config']['unconverged_nodes'])
except Exception as ex: print('failed',ex)
PY
