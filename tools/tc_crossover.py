#!/usr/bin/env python
"""Crossover table of the two stage-1 implementations behind `ancuts_affinity_f32` (dense N x N float32 output): the
pair-queue kernels (float64 distance test on all pairs, feature distances for the in-mask pairs only) against the
tcgen05 Gram GEMM (`affinity_impl = 1`: all N^2 dot products, 3 x TF32), for N in 2 k .. 16 k and F in {96, 480}.

    python tools/tc_crossover.py --out gpurun_out/tc_crossover.json [--only-tc N]     (--only-tc: one tcgen05 launch, for ncu)
"""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from autoinst_b200 import api
from autoinst_b200.synthetic import make_chunk

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="gpurun_out/tc_crossover.json")
ap.add_argument("--only-tc", dest="only_tc", type=int, default=0)
args = ap.parse_args()
dev = torch.device("cuda", 0)


def inputs(nt, dino):
    ch = make_chunk(900 + nt, n_target=nt, features="tarl_dino" if dino else "tarl")
    return (ch.n, torch.as_tensor(ch.points, device=dev), torch.as_tensor(ch.tarl, dtype=torch.float32, device=dev),
            torch.as_tensor(ch.dino, dtype=torch.float32, device=dev) if dino else None)


if args.only_tc:
    n, pts, tarl, dino = inputs(args.only_tc, False)
    for _ in range(3):
        api.affinity(pts, tarl, alpha=1.0, theta=0.5, device=dev, impl=1)
    torch.cuda.synchronize()
    print("ok", n)
    sys.exit(0)

rows = []
for nt in (2048, 4096, 8192, 16384):
    for use_dino in (False, True):
        n, pts, tarl, dino = inputs(nt, use_dino)
        F = 96 + (384 if use_dino else 0)
        kw = dict(alpha=1.0, theta=0.5, gamma=0.1 if use_dino else 0.0, device=dev)
        ms = {}
        for impl in (0, 1):
            for _ in range(3):
                W = api.affinity(pts, tarl, dino, impl=impl, **kw)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                W = api.affinity(pts, tarl, dino, impl=impl, **kw)
            b.record(); torch.cuda.synchronize()
            ms[impl] = a.elapsed_time(b) / 10
        nnz = int((W != 0).sum().item())
        rows.append(dict(n=n, F=F, pair_kernels_ms=ms[0], tcgen05_ms=ms[1], tc_over_pair=ms[1] / ms[0], in_mask_fraction=nnz / (n * n),
                         gemm_tflops_2n2f=2.0 * n * n * F / ms[1] / 1e9, mma_tflops_3xtf32=6.0 * n * n * F / ms[1] / 1e9,
                         dense_store_gbs_pair=4.0 * n * n / ms[0] / 1e6))
        print(json.dumps(rows[-1]), flush=True)
os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
json.dump(rows, open(args.out, "w"), indent=1)
