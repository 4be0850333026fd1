#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 --cpu-chunks 0 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N exit $?" > gpurun_out/summary_multi.txt
tail -c 600 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 tools/map_eval.py --chunks 24 --n-per-chunk 6000 --no-oracle --out gpurun_out/map_eval_n$N.json > gpurun_out/map_eval_n$N.log 2>&1; echo "map N=$N exit $?" >> gpurun_out/summary_multi.txt
tail -1 gpurun_out/map_eval_n$N.log | cut -c1-600
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 --ref-chunks 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?" >> gpurun_out/summary_multi.txt
cat gpurun_out/bench_ref.json | cut -c1-700
cat gpurun_out/summary_multi.txt
