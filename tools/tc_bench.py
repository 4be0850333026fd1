#!/usr/bin/env python
"""Time the two affinity implementations (exact tile kernel vs tcgen05 Gram GEMM) with device-resident inputs."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from autoinst_b200 import api
from autoinst_b200.synthetic import make_chunk

dev = torch.device("cuda", 0)
rows = []
for nt in (4096, 8192, 16384):
    ch = make_chunk(900 + nt, n_target=nt, features="tarl")
    pts = torch.as_tensor(ch.points, device=dev)
    tarl = torch.as_tensor(ch.tarl, dtype=torch.float32, device=dev)
    n = ch.n
    for impl in (0, 1):
        for _ in range(3):
            W = api.affinity(pts, tarl, alpha=1.0, theta=0.5, device=dev, impl=impl)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            W = api.affinity(pts, tarl, alpha=1.0, theta=0.5, device=dev, impl=impl)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        rows.append(dict(n=n, impl=impl, ms=ms, store_gbs=4.0 * n * n / ms / 1e6, gemm_tflops=2.0 * n * n * 96 / ms / 1e9,
                         mma_tflops_3xtf32=3 * 2.0 * n * n * 96 / ms / 1e9))
        print(json.dumps(rows[-1]), flush=True)
json.dump(rows, open("gpurun_out/tc_bench.json", "w"), indent=1)
