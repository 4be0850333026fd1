#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_stages.py -m gpu -q -x --tb=short -p no:cacheprovider -k "lanczos" > gpurun_out/tests_lanczos.log 2>&1; echo "lanczos tests exit $?" > gpurun_out/summary.txt
tail -15 gpurun_out/tests_lanczos.log
timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt
tail -3 gpurun_out/tests.log
timeout 900 python bench.py --steps 3 --warmup 3 --cpu-chunks 0 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'])
print('stage_ms',d['config']['stage_ms_one_step'])
print('roofline',d['roofline']['achieved'],d['roofline']['frac'],d['roofline']['avg_launch_us'])
print('steps/chunk',d['config']['lanczos_steps_per_chunk'],'unconv',d['config']['unconverged_nodes'])
PY
tail -3 gpurun_out/bench.err
cat gpurun_out/summary.txt
