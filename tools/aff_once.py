#!/usr/bin/env python
"""One stage-1 call on a synthetic chunk (profiling target for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from autoinst_b200 import api
from autoinst_b200.synthetic import make_chunk
dev = torch.device("cuda", 0)
ch = make_chunk(900 + 8192, n_target=8192, features="tarl")
pts = torch.as_tensor(ch.points, device=dev)
tarl = torch.as_tensor(ch.tarl, dtype=torch.float32, device=dev)
for _ in range(3):
    W = api.affinity(pts, tarl, alpha=1.0, theta=0.5, device=dev)
torch.cuda.synchronize()
print("ok", ch.n)
