#!/usr/bin/env python
"""BASELINE.json config 5: affinity + eigensolve scaling sweep over the chunk size N on one B200,
with per-kernel HBM figures for nodes large enough to be HBM-bound (one connected "street" scene per N).

    python tools/nsweep.py --sizes 4096 8192 16384 32768 --out gpurun_out/nsweep.json [--cpu]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def street_scene(n_target, seed=0):
    """One connected component: a dense slab of voxel centres (jittered grid) sized to n_target points."""
    rng = np.random.default_rng(seed)
    nx = int(round((n_target / 4) ** 0.5 * 1.6))
    ny = max(2, -(-n_target // (4 * nx)))
    g = np.stack(np.meshgrid(np.arange(nx), np.arange(ny), np.arange(4), indexing="ij"), -1).reshape(-1, 3) * 0.35
    g = g + rng.uniform(-0.1, 0.1, size=g.shape)
    g = g[rng.permutation(len(g))[:n_target]]
    return g.astype(np.float32).astype(np.float64)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", type=int, nargs="+", default=[4096, 8192, 16384, 32768])
    ap.add_argument("--out", default="gpurun_out/nsweep.json")
    ap.add_argument("--cpu", action="store_true", help="also time scipy cdist + eigsh (top level only) on the host")
    ap.add_argument("--max-steps", type=int, default=64)
    args = ap.parse_args()
    import torch
    from autoinst_b200 import api
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    dev = torch.device("cuda", 0)
    hd = api.Handle.get(dev)
    rows = []
    for n in args.sizes:
        pts = street_scene(n, seed=n)
        n = pts.shape[0]
        rng = np.random.default_rng(n)
        tarl = rng.normal(size=(n, 96)).astype(np.float32)
        ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pts_np, tarl_np = pts, tarl
        pts = torch.as_tensor(pts_np, device=dev)                 # inputs resident in HBM: time the kernels, not the copies
        tarl = torch.as_tensor(tarl_np, dtype=torch.float32, device=dev)
        # stage 1
        for _ in range(2):
            W = api.affinity(pts, tarl, alpha=1.0, theta=0.5, device=dev)
        torch.cuda.synchronize()
        ev_a.record()
        for _ in range(3):
            W = api.affinity(pts, tarl, alpha=1.0, theta=0.5, device=dev)
        ev_b.record(); torch.cuda.synchronize()
        t_aff = ev_a.elapsed_time(ev_b) / 3
        # stage 2
        for _ in range(2):
            deg = api.degree_normalize(W)
        torch.cuda.synchronize(); ev_a.record()
        for _ in range(5):
            deg = api.degree_normalize(W)
        ev_b.record(); torch.cuda.synchronize()
        t_deg = ev_a.elapsed_time(ev_b) / 5
        for _ in range(3):              # the output matrix is allocated per call: let the caching allocator settle first
            deg, M = api.degree_normalize(W, return_normalized=True)
        torch.cuda.synchronize(); ev_a.record()
        for _ in range(5):
            deg, M = api.degree_normalize(W, return_normalized=True)
        ev_b.record(); torch.cuda.synchronize()
        t_norm = ev_a.elapsed_time(ev_b) / 5
        del M
        # stage 3: grid-wide matvec path on the whole matrix as ONE node, fixed number of steps
        hd.set_stage_timing(1)
        ev, lam2, steps, conv = api.lanczos_fiedler(W, [0], [n], max_steps=args.max_steps, lanczos_impl=1)
        acc = hd.accounting()
        hd.set_stage_timing(0)
        mv = acc["matvec"]
        row = dict(n=n, nnz_per_row=float((W != 0).sum().item()) / n,
                   affinity_ms=t_aff, affinity_gbs=(4.0 * n * n + 4.0 * n * 99) / t_aff / 1e6,
                   degree_ms=t_deg, degree_gbs=4.0 * n * n / t_deg / 1e6,
                   normalize_ms=t_norm, normalize_gbs=12.0 * n * n / t_norm / 1e6,
                   matvec_launches=mv["launches"], matvec_avg_us=1e3 * mv["ms"] / max(mv["launches"], 1),
                   matvec_gbs=mv["bytes"] / max(mv["ms"], 1e-9) / 1e6, reorth_ms_per_step=acc["reorth"]["ms"] / max(int(steps[0]), 1),
                   lanczos_steps=int(steps[0]), lambda2=float(lam2[0]), converged=int(conv[0]), hbm_peak_gbs=peak)
        for k in ("affinity", "degree", "normalize", "matvec"):
            row[k + "_frac_of_hbm_peak"] = row[k + "_gbs"] / peak
        if args.cpu and n <= 16384:
            from scipy.spatial.distance import cdist
            import scipy.sparse as sp
            from oracle.affinity_ref import affinity_ref
            from oracle.ncut_ref import fiedler_of_block
            t0 = time.perf_counter()
            A = affinity_ref(pts_np, tarl_np.astype(np.float64), alpha=1.0, theta=0.5)
            t1 = time.perf_counter()
            d, D, evr, vals = fiedler_of_block(sp.csr_matrix(A))
            t2 = time.perf_counter()
            row.update(cpu_affinity_s=t1 - t0, cpu_eigsh_s=t2 - t1, cpu_lambda2=float(vals[1]))
        rows.append(row)
        print(json.dumps(row), flush=True)
        del W
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
