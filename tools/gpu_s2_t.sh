#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt
tail -3 gpurun_out/tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --cpu-chunks 0 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
sm=d['config']['stage_ms_one_step']
print('value %.1f'%d['value'],'ms %.2f'%d['ms_per_step'],'e2e %.1f ms %.2f'%(d['e2e']['value'],d['e2e']['ms_per_step']),'aff %.2f mv %.2f part %.2f'%(sm['affinity'],sm['matvec'],sm['partition']),'frac %.3f'%d['roofline']['frac'],'traffic',d['roofline']['traffic'])
PY
cat gpurun_out/summary.txt
