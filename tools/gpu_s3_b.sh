#!/bin/bash
# session 3, run B: cheaper serial part of the convergence check (parallel Gershgorin bounds, reciprocal pivots), check split clock
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/tests_s3b.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt
tail -3 gpurun_out/tests_s3b.log
ANCUTS_PHASES=1 timeout 400 python tools/level_profile.py --batch 128 --out gpurun_out/levels_s3b.json > gpurun_out/levels_s3b.log 2>&1; echo "levels exit $?" >> gpurun_out/summary.txt
grep "cluster size" gpurun_out/levels_s3b.log
timeout 600 python bench.py --steps 3 --warmup 3 --cpu-chunks 0 --batch 128 > gpurun_out/bench_s3b.json 2> gpurun_out/bench_s3b.err; echo "bench exit $?" >> gpurun_out/summary.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_s3b.json'))
sm=d['config']['stage_ms_one_step']
print('value %.1f'%d['value'],'ms %.2f'%d['ms_per_step'],'e2e %.1f'%d['e2e']['value'],'aff %.2f mv %.2f part %.2f'%(sm['affinity'],sm['matvec'],sm['partition']),'frac %.3f'%d['roofline']['frac'],'steps',d['config']['lanczos_steps_per_chunk'],'unconv',d['config']['unconverged_nodes'])
PY
timeout 900 python tools/parity_sweep.py --config tarl_spatial --chunks 32 --n-target 8192 --seed 7000 --oracle-cache parity_cache --out gpurun_out/parity_tarl_spatial.json > gpurun_out/parity_tarl_spatial.log 2>&1; echo "parity exit $?" >> gpurun_out/summary.txt
tail -1 gpurun_out/parity_tarl_spatial.log | cut -c1-250
cat gpurun_out/summary.txt
