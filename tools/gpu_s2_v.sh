#!/bin/bash
# session 2, run V: 1024-thread CTAs (64 registers, 8 warps per scheduler, 2 KB ring stages) vs 512-thread default
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
export T1024=$PWD/autoinst_b200/lib/libautoinst_ncuts_t1024.so
ANCUTS_LIB_PATH=$T1024 timeout 600 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/tests_t1024.log 2>&1; echo "tests t1024 exit $?" >> gpurun_out/summary.txt
tail -3 gpurun_out/tests_t1024.log
run() {
  ANCUTS_LIB_PATH=$2 timeout 600 python bench.py --steps 3 --warmup 3 --cpu-chunks 0 --batch $1 > gpurun_out/bench_$3_b$1.json 2> gpurun_out/bench_$3_b$1.err; echo "bench $3 b$1 exit $?" >> gpurun_out/summary.txt
  python - $1 $3 <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/bench_%s_b%s.json'%(sys.argv[2],sys.argv[1])))
    sm=d['config']['stage_ms_one_step']
    print(sys.argv[2],'batch',sys.argv[1],'value %.1f'%d['value'],'ms %.2f'%d['ms_per_step'],'e2e %.1f'%d['e2e']['value'],'aff %.2f mv %.2f part %.2f'%(sm['affinity'],sm['matvec'],sm['partition']),'frac %.3f'%d['roofline']['frac'],'steps',d['config']['lanczos_steps_per_chunk'],'seg',d['config']['segments_per_chunk'],'unconv',d['config']['unconverged_nodes'])
except Exception as ex: print('failed',sys.argv[1:],ex)
PY
}
run 128 "" t512; run 128 $T1024 t1024; run 64 $T1024 t1024; run 16 $T1024 t1024; run 16 "" t512
cat gpurun_out/summary.txt
