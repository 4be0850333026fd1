#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/nsweep.py --sizes 4096 8192 16384 32768 --out gpurun_out/nsweep.json > gpurun_out/nsweep.log 2>&1; echo "sweep exit $?"
python - <<'PY'
import json
for r in json.load(open('gpurun_out/nsweep.json')):
    print({k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items() if k in ('n','affinity_ms','affinity_frac_of_hbm_peak','degree_frac_of_hbm_peak','normalize_frac_of_hbm_peak','matvec_avg_us','matvec_frac_of_hbm_peak','reorth_ms_per_step','lanczos_steps','converged','nnz_per_row')})
PY
tail -3 gpurun_out/nsweep.log
