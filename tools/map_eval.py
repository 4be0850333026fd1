#!/usr/bin/env python
"""BASELINE.json config 4: full synthetic first-map NCuts pass, chunks sharded over the ranks, labels gathered
over NCCL, merged and scored; the same through the CPU oracle for parity level 3.

    python tools/map_eval.py --chunks 40 --n-per-chunk 8192 --out gpurun_out/map_eval.json        (1 GPU)
    torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/map_eval.py ...                      (N GPUs)
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def oracle_job(a):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    pts, tarl, cfg = a
    import scipy.sparse as sp
    from oracle import ncut_ref as R
    from oracle.affinity_ref import affinity_ref
    t0 = time.perf_counter()
    A = affinity_ref(pts, tarl, None, alpha=cfg["alpha"], theta=cfg["theta"], gamma=cfg["gamma"])
    with R.pinned_eigsh():
        g = R.normalized_cut_ref(sp.csr_matrix(A), pts.shape[0], np.arange(pts.shape[0]), T=cfg["T"])
    return R.labels_from_groups(g, pts.shape[0]), time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=40)
    ap.add_argument("--n-per-chunk", type=int, default=8192)
    ap.add_argument("--seed", type=int, default=77)
    ap.add_argument("--no-oracle", action="store_true")
    ap.add_argument("--out", default="gpurun_out/map_eval.json")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from autoinst_b200 import api, sharding
    from autoinst_b200.synthetic import CONFIGS, make_map
    from oracle import merge_ref as M
    from oracle import ncut_ref as R
    from oracle.metrics_ref import instance_metrics
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = CONFIGS["tarl_spatial"]
    chunks = make_map(args.chunks, args.n_per_chunk, seed=args.seed)
    sizes = [c.n for c in chunks]
    mine = sharding.shard_chunks(sizes, world)[rank]
    kw = dict(alpha=cfg["alpha"], theta=cfg["theta"], T=cfg["T"], device=dev)
    api.segment_chunks([chunks[i].points for i in mine[:1]], [chunks[i].tarl for i in mine[:1]], **kw)    # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = api.segment_chunks([chunks[i].points for i in mine], [chunks[i].tarl for i in mine], **kw)
    labels = sharding.gather_labels(mine, res.labels, len(chunks), device=dev)
    torch.cuda.synchronize()
    t_gpu = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([t_gpu], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_gpu = float(t.item())
    if rank == 0:
        gt_pts, gt_lab = M.merge_unite_gt([(c.points, c.instance) for c in chunks])
        gt = M.compact_labels(gt_lab)

        def evaluate(per_chunk):
            parts = [(c.points, M.globally_unique(c.chunk_id, M.canonical_labels(lab))) for c, lab in zip(chunks, per_chunk)]
            pts, lab = M.merge_chunks_unite_instances(parts)
            allp = M.compact_labels(lab)
            return instance_metrics(allp, M.remove_semantics(gt, allp.copy()), gt, min_points=20)

        out = dict(chunks=len(chunks), points=int(sum(sizes)), n_gpus=world, gpu_seconds=t_gpu,
                   chunks_per_sec=len(chunks) / t_gpu, metrics_gpu=evaluate(labels))
        if not args.no_oracle:
            t1 = time.perf_counter()
            workers = min(os.cpu_count() or 1, 32, len(chunks))
            with mp.get_context("spawn").Pool(workers) as pool:
                o = pool.map(oracle_job, [(c.points, c.tarl, cfg) for c in chunks])
            t_cpu = time.perf_counter() - t1
            ref_labels = [x[0] for x in o]
            out.update(metrics_oracle=evaluate(ref_labels), oracle_wall_seconds=t_cpu, oracle_workers=workers,
                       oracle_cpu_seconds_sum=float(sum(x[1] for x in o)),
                       chunks_identical=int(sum(R.same_partition(a, b) for a, b in zip(labels, ref_labels))),
                       max_metric_diff=None)
            out["max_metric_diff"] = max(abs(out["metrics_gpu"][k] - out["metrics_oracle"][k]) for k in out["metrics_gpu"])
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        json.dump(out, open(args.out, "w"), indent=1)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
