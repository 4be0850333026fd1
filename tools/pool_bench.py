#!/usr/bin/env python
"""Feature pooling row (N3): GPU time of ancuts_feature_pool (CUDA events, inputs resident and through host buffers)
next to the CPU oracle (the reference's KD-tree loop restated), on one synthetic chunk with 21 scans.

    python tools/pool_bench.py --n-target 8192 --out gpurun_out/pool_bench.json
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from autoinst_b200 import api
from autoinst_b200.synthetic import make_chunk, make_scans
from oracle.pool_ref import pool_features_ref

ap = argparse.ArgumentParser()
ap.add_argument("--n-target", type=int, default=8192)
ap.add_argument("--scans", type=int, default=21)
ap.add_argument("--pts-per-major", type=float, default=3.0)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--out", default="gpurun_out/pool_bench.json")
args = ap.parse_args()
dev = torch.device("cuda", 0)
ch = make_chunk(3, n_target=args.n_target, features="tarl")
scans = make_scans(ch, n_scans=args.scans, pts_per_major=args.pts_per_major)
pts = np.concatenate([s[0] for s in scans]); fts = np.concatenate([s[1] for s in scans])
lo, hi = ch.center - 12.5, ch.center + 12.5
R = 0.175
t0 = time.time(); ref, rc = pool_features_ref(ch.points, scans, ch.center, radius=R, return_count=True); t_cpu = time.time() - t0
dm = torch.as_tensor(ch.points, device=dev); dp = torch.as_tensor(pts, device=dev); df = torch.as_tensor(fts, device=dev)
for _ in range(3):
    out, cnt = api.feature_pool(dm, dp, df, R, lo, hi, return_count=True, device=dev)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(args.iters):
    out, cnt = api.feature_pool(dm, dp, df, R, lo, hi, return_count=True, device=dev)
b.record(); torch.cuda.synchronize()
ms_res = a.elapsed_time(b) / args.iters
hp = torch.as_tensor(pts).pin_memory(); hf = torch.as_tensor(fts).pin_memory(); hm = torch.as_tensor(ch.points).pin_memory()
torch.cuda.synchronize(); t0 = time.time()
for _ in range(5):
    o = api.feature_pool(hm.to(dev, non_blocking=True), hp.to(dev, non_blocking=True), hf.to(dev, non_blocking=True), R, lo, hi, device=dev).cpu()
ms_e2e = (time.time() - t0) / 5 * 1e3
out = out.cpu().numpy(); cnt = cnt.cpu().numpy()
hits = int(cnt.sum()); m = pts.shape[0]; n = ch.n; F = fts.shape[1]
alg_bytes = m * 24 + hits * (24 + 4 * F) + n * (24 + 8 * F)     # every scan point read once for its key, hit rows once, output once
res = dict(n_major=n, n_scan=m, feat_dim=F, hits=hits, zero_rows=int((cnt == 0).sum()),
           counts_equal=bool(np.array_equal(cnt, rc)), max_abs_err=float(np.abs(out - ref).max()),
           gpu_ms_resident=ms_res, gpu_ms_host_buffers=ms_e2e, cpu_oracle_s=t_cpu, cpu_cores=1,
           speedup_resident=t_cpu * 1e3 / ms_res, speedup_host_buffers=t_cpu * 1e3 / ms_e2e,
           algorithmic_mb=alg_bytes / 1e6, achieved_gbs=alg_bytes / 1e9 / (ms_res * 1e-3))
os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
json.dump(res, open(args.out, "w"), indent=1)
print(json.dumps(res))
