#!/usr/bin/env python
"""One warm-up pass and one pass of the bench workload with inputs resident in HBM (profiling target for ncu).

    python tools/one_step.py --batch 128
"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from autoinst_b200 import api
from autoinst_b200.synthetic import CONFIGS, make_chunk
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--n-target", dest="n_target", type=int, default=8192)
ap.add_argument("--seed", type=int, default=1000)
ap.add_argument("--passes", type=int, default=2)
ap.add_argument("--matvec", type=int, default=0, help="ANCUTS_OPT_MATVEC: 0 shared-memory sparse (default), 1 dense from HBM")
ap.add_argument("--pairs", type=int, default=0, help="ANCUTS_OPT_PAIR_SEARCH: 0 cell-sorted (default), 1 shuffled")
args = ap.parse_args()
dev = torch.device("cuda", 0)
cfg = CONFIGS["tarl_spatial"]
chunks = [make_chunk(args.seed + i, n_target=args.n_target, features="tarl") for i in range(args.batch)]
packed = api.PackedChunks([c.points for c in chunks], [c.tarl for c in chunks], None, theta=cfg["theta"])
devc = packed.to_device(dev)
from autoinst_b200._lib import OPT_MATVEC, OPT_PAIR_SEARCH
api.Handle.get(dev).set_option(OPT_MATVEC, args.matvec)
api.Handle.get(dev).set_option(OPT_PAIR_SEARCH, args.pairs)
for _ in range(args.passes):
    api.segment_packed(packed, dev_chunks=devc, alpha=cfg["alpha"], theta=cfg["theta"], T=cfg["T"])
torch.cuda.synchronize()
print("ok", args.batch, int(packed.off[-1]))
