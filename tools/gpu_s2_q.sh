#!/bin/bash
# session 2, run Q: branch-free ring stages; N sweep with resident inputs and the new grid-matvec dispatch
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt
tail -3 gpurun_out/tests.log
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
timeout 600 python tools/nsweep.py --sizes 4096 8192 16384 32768 --out gpurun_out/nsweep.json > gpurun_out/nsweep.log 2>&1; echo "nsweep exit $?" >> gpurun_out/summary.txt
python - <<'PY'
import json
for r in json.load(open('gpurun_out/nsweep.json')):
    print('n',r['n'],'aff %.3f ms %.2f'%(r['affinity_ms'],r['affinity_frac_of_hbm_peak']),'deg %.2f'%r['degree_frac_of_hbm_peak'],'norm %.2f'%r['normalize_frac_of_hbm_peak'],'matvec %.2f'%r['matvec_frac_of_hbm_peak'])
PY
cat gpurun_out/summary.txt
