#!/bin/bash
# One parameterised runner for the GPU box (replaces the per-session scripts of round 1):
#   gpurun --timeout 2400 -- 'bash tools/gpu_run.sh tests bench small map levels refarm ncu_list'
# Every stage writes under gpurun_out/<tag>_*; a failing stage does not stop the later ones.
cd "$(dirname "$0")/.." || exit 1
TAG=${TAG:-r2}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for stage in "$@"; do
  echo "=== stage $stage $(date +%T)"
  case $stage in
    tests)    timeout 1500 python -m pytest tests -x -q -m gpu --durations=15 > $OUT/${TAG}_tests.log 2>&1; echo "rc=$?" >> $OUT/${TAG}_tests.log; tail -5 $OUT/${TAG}_tests.log ;;
    testsall) timeout 1500 python -m pytest tests -q -m gpu --durations=15 > $OUT/${TAG}_tests.log 2>&1; echo "rc=$?" >> $OUT/${TAG}_tests.log; tail -15 $OUT/${TAG}_tests.log ;;
    smoke)    timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -2 $OUT/${TAG}_smoke.log ;;
    bench)    timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 600 $OUT/${TAG}_bench.json ;;
    bench20)  timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench20.json 2> $OUT/${TAG}_bench20.err; tail -c 400 $OUT/${TAG}_bench20.json ;;
    benchdense) timeout 900 python bench.py --steps 5 --warmup 3 --matvec dense --cpu-chunks 0 > $OUT/${TAG}_bench_dense.json 2> $OUT/${TAG}_bench_dense.err; tail -c 400 $OUT/${TAG}_bench_dense.json ;;
    ab)       for v in "shuffled dense" "sorted dense" "shuffled sparse" "sorted sparse"; do set -- $v; timeout 600 python bench.py --steps 4 --warmup 3 --pairs $1 --matvec $2 --cpu-chunks 0 --no-python-surface --roof-steps 0 > $OUT/${TAG}_ab_$1_$2.json 2> $OUT/${TAG}_ab_$1_$2.err; python -c "import json;d=json.load(open('$OUT/${TAG}_ab_$1_$2.json'));print('pairs','$1','matvec','$2',round(d['value'],1),round(d['e2e']['value'],1),d['roofline']['frac'],d['detail']['stage_ms_one_step'])"; done ;;
    libab)    for v in ${LIBS:-default r4 hints r4hints}; do L=""; [ $v != default ] && L="autoinst_b200/lib_ab/libautoinst_ncuts_$v.so"; for mv in ${LIBAB_MV:-dense}; do ANCUTS_LIB_PATH=$L timeout 600 python bench.py --steps 4 --warmup 3 --matvec $mv --cpu-chunks 0 --no-python-surface --roof-steps 0 > $OUT/${TAG}_lib_${v}_$mv.json 2> $OUT/${TAG}_lib_${v}_$mv.err; python -c "import json;d=json.load(open('$OUT/${TAG}_lib_${v}_$mv.json'));print('lib','$v','matvec','$mv',round(d['value'],1),round(d['e2e']['value'],1),d['roofline']['frac'],d['detail']['stage_ms_one_step']['matvec'])"; done; done ;;
    tc)       timeout 600 python tools/tc_crossover.py --out $OUT/${TAG}_tc_crossover.json > $OUT/${TAG}_tc_crossover.log 2>&1; tail -8 $OUT/${TAG}_tc_crossover.log | cut -c1-250; \
              timeout 200 python tools/tc_crossover.py --only-tc 8192 > $OUT/${TAG}_tc_plain.log 2>&1 && \
              timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_affinity_tc -s 2 -c 1 -o $OUT/${TAG}_prof_affinity_tc -f python tools/tc_crossover.py --only-tc 8192 > $OUT/${TAG}_tc_ncu.log 2>&1; tail -2 $OUT/${TAG}_tc_ncu.log ;;
    cmaps)    for m in ${CMAPS:-122488 112488 112248 122448 124488}; do timeout 600 python bench.py --steps 4 --warmup 3 --cluster-map $m --cpu-chunks 0 --no-python-surface --roof-steps 0 > $OUT/${TAG}_cmap_$m.json 2> $OUT/${TAG}_cmap_$m.err; python -c "import json;d=json.load(open('$OUT/${TAG}_cmap_$m.json'));print('cmap',$m,round(d['value'],1),round(d['e2e']['value'],1),d['detail']['stage_ms_one_step']['matvec'])"; done ;;
    configs)  for c in spatial tarl_spatial_dino; do timeout 900 python bench.py --steps 3 --warmup 3 --config $c --cpu-chunks 0 > $OUT/${TAG}_bench_$c.json 2> $OUT/${TAG}_bench_$c.err; tail -c 300 $OUT/${TAG}_bench_$c.json; done ;;
    small)    for b in 40 16 5 1; do timeout 600 python bench.py --steps 5 --warmup 3 --batch $b --cpu-chunks 0 --no-python-surface > $OUT/${TAG}_bench_b$b.json 2> $OUT/${TAG}_bench_b$b.err; python -c "import json;d=json.load(open('$OUT/${TAG}_bench_b$b.json'));print('batch',$b,d['value'],d['e2e']['value'],d['roofline']['frac'])"; done ;;
    map)      timeout 900 python bench.py --workload map --steps 3 --warmup 2 > $OUT/${TAG}_bench_map.json 2> $OUT/${TAG}_bench_map.err; tail -c 900 $OUT/${TAG}_bench_map.json ;;
    levels)   ANCUTS_PHASES=1 timeout 600 python tools/level_profile.py --batch 128 --matvec 1 --out $OUT/${TAG}_levels_b128.json > $OUT/${TAG}_levels.log 2>&1; grep "cluster size" $OUT/${TAG}_levels.log ;;
    levelsp)  ANCUTS_PHASES=1 timeout 600 python tools/level_profile.py --batch 128 --matvec 0 --pairs 0 --out $OUT/${TAG}_levels_b128_sparse.json > $OUT/${TAG}_levels_sparse.log 2>&1; grep "cluster size" $OUT/${TAG}_levels_sparse.log; python -c "import json;d=json.load(open('$OUT/${TAG}_levels_b128_sparse.json'));print([(l['active'],round(l['ms'],2)) for l in d['levels']])" ;;
    scale)    N=${NGPU:-2}; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --cpu-chunks 0 --no-python-surface > $OUT/${TAG}_scale_n$N.json 2> $OUT/${TAG}_scale_n$N.err; python -c "import json;d=json.load(open('$OUT/${TAG}_scale_n$N.json'));print('N',$N,round(d['value'],1),round(d['e2e']['value'],1),d['ms_per_step'],d['detail']['ms_per_step_over_ranks'],d['detail']['gather_ms'])" || tail -5 $OUT/${TAG}_scale_n$N.err ;;
    scalemap) N=${NGPU:-2}; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload map --steps 5 --warmup 3 > $OUT/${TAG}_scalemap_n$N.json 2> $OUT/${TAG}_scalemap_n$N.err; python -c "import json;d=json.load(open('$OUT/${TAG}_scalemap_n$N.json'));print('map N',$N,round(d['value'],1),round(d['e2e']['value'],1),d['ms_per_step'],d['detail']['chunks_per_rank'],d['detail']['metrics'])" || tail -5 $OUT/${TAG}_scalemap_n$N.err ;;
    refarm)   timeout 1200 python bench.py --impl reference --steps ${REF_STEPS:-4} --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; tail -c 1200 $OUT/${TAG}_bench_reference.json ;;
    ncu_list) timeout 300 python tools/one_step.py --batch 128 --passes 2 --matvec ${ONE_STEP_MATVEC:-1} > $OUT/${TAG}_one_step.log 2>&1 && \
              timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv \
                  --log-file $OUT/${TAG}_launches_raw.csv python tools/one_step.py --batch 128 --passes 2 --matvec ${ONE_STEP_MATVEC:-1} > $OUT/${TAG}_ncu_list.log 2>&1; \
              python tools/ncu_summarise.py $OUT/${TAG}_launches_raw.csv > $OUT/${TAG}_launch_list_b128.csv 2>> $OUT/${TAG}_ncu_list.log; head -30 $OUT/${TAG}_launch_list_b128.csv ;;
    ncu_full) timeout 300 python tools/one_step.py --batch 128 --passes 1 --matvec ${ONE_STEP_MATVEC:-1} > $OUT/${TAG}_one_step.log 2>&1 && \
              timeout 900 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-k_lanczos_cluster} -s ${NCU_SKIP:-1} -c ${NCU_COUNT:-2} \
                  -o $OUT/${TAG}_prof_${NCU_NAME:-cluster} -f python tools/one_step.py --batch 128 --passes 1 --matvec ${ONE_STEP_MATVEC:-1} > $OUT/${TAG}_ncu_full.log 2>&1; tail -3 $OUT/${TAG}_ncu_full.log ;;
    parity)   for c in spatial tarl_spatial tarl_spatial_dino; do timeout 900 python tools/parity_sweep.py --config $c --chunks 32 --n-target 8192 --seed 7000 --oracle-cache parity_cache --out $OUT/${TAG}_parity_$c.json > $OUT/${TAG}_parity_$c.log 2>&1; tail -c 400 $OUT/${TAG}_parity_$c.log; done ;;
    *) echo "unknown stage $stage" ;;
  esac
done
echo "=== done $(date +%T)"
