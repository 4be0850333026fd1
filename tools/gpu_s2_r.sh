#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt
tail -3 gpurun_out/tests.log
timeout 600 python tools/nsweep.py --sizes 4096 8192 16384 32768 --out gpurun_out/nsweep.json > gpurun_out/nsweep.log 2>&1; echo "nsweep exit $?" >> gpurun_out/summary.txt
python - <<'PY'
import json
for r in json.load(open('gpurun_out/nsweep.json')):
    print('n',r['n'],'aff %.3f ms %.2f'%(r['affinity_ms'],r['affinity_frac_of_hbm_peak']),'deg %.2f'%r['degree_frac_of_hbm_peak'],'norm %.3f ms %.2f'%(r['normalize_ms'],r['normalize_frac_of_hbm_peak']),'matvec %.2f'%r['matvec_frac_of_hbm_peak'],'steps',r['lanczos_steps'])
PY
cat gpurun_out/summary.txt
