#!/usr/bin/env python
"""Throughput of one batch vs the same chunks split over 2 / 4 concurrent lanes (handle + stream + host thread each).

    python tools/lanes_bench.py --chunks 64
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from autoinst_b200 import api
from autoinst_b200.synthetic import CONFIGS, make_chunk
ap = argparse.ArgumentParser()
ap.add_argument("--chunks", type=int, default=16)
ap.add_argument("--lanes", type=int, nargs="+", default=[1, 2, 4])
args = ap.parse_args()
dev = torch.device("cuda", 0)
cfg = CONFIGS["tarl_spatial"]
kw = dict(alpha=cfg["alpha"], theta=cfg["theta"], T=cfg["T"])
chunks = [make_chunk(1000 + i, n_target=8192, features="tarl") for i in range(args.chunks)]
for lanes in args.lanes:
    parts = [chunks[i::lanes] for i in range(lanes)]
    packed = [api.PackedChunks([c.points for c in p], [c.tarl for c in p], theta=cfg["theta"]) for p in parts]
    devs = [pk.to_device(dev) for pk in packed]
    for _ in range(3):
        api.segment_packed_lanes(packed, devs, **kw)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        api.segment_packed_lanes(packed, devs, **kw)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(json.dumps(dict(chunks=args.chunks, lanes=lanes, ms_per_pass=ms, chunks_per_sec=args.chunks / ms * 1e3)), flush=True)
