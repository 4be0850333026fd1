#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "tests exit $?" > gpurun_out/summary.txt
tail -3 gpurun_out/tests.log
timeout 600 python tools/nsweep.py --sizes 8192 32768 --out gpurun_out/nsweep.json > gpurun_out/nsweep.log 2>&1; echo "sweep exit $?" >> gpurun_out/summary.txt
python - <<'PY'
import json
for r in json.load(open('gpurun_out/nsweep.json')):
    print({k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items() if k in ('n','affinity_ms','affinity_frac_of_hbm_peak','degree_frac_of_hbm_peak','normalize_frac_of_hbm_peak','matvec_frac_of_hbm_peak')})
PY
timeout 900 python bench.py --steps 3 --warmup 3 --cpu-chunks 0 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'])
print('stage_ms',d['config']['stage_ms_one_step'])
PY
timeout 900 python tools/map_eval.py --chunks 24 --n-per-chunk 6000 --out gpurun_out/map_eval.json > gpurun_out/map_eval.log 2>&1; echo "map exit $?" >> gpurun_out/summary.txt
tail -2 gpurun_out/map_eval.log | cut -c1-1500
cat gpurun_out/summary.txt
