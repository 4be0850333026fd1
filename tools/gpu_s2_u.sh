#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
ANCUTS_PHASES=1 timeout 400 python tools/level_profile.py --batch 128 --out gpurun_out/levels_b128.json > gpurun_out/levels_b128.log 2>&1; echo "levels exit $?" >> gpurun_out/summary.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/levels_b128.json'))
for l in d['levels']:
    print(l['level'],l['active'],l['bins'],l['cluster'],'ms %.2f'%l['ms'],'maxn',l['max_n'],'maxsteps',l['max_steps'],'mean %.0f'%l['mean_steps'],'GB %.1f'%l['gb'],'GB/s %.0f'%l['gbs'])
print('matvec ms',d['matvec_ms'])
PY
grep "cluster size" gpurun_out/levels_b128.log
timeout 600 python tools/lanes_bench.py --chunks 128 --lanes 1 2 4 > gpurun_out/lanes128.log 2>&1; echo "lanes exit $?" >> gpurun_out/summary.txt
cat gpurun_out/lanes128.log | tail -3
cat gpurun_out/summary.txt
