"""Does splitting a small batch over concurrent lanes (api.segment_packed_lanes: one handle + stream + host thread each) pay?
One synthetic 40-chunk map (or --batch chunks of --n-target points), inputs resident, host timer around the whole call.
    python tools/lanes_probe.py [--map 40 | --batch 16 --n-target 8192] [--lanes 1 2 4]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--map", type=int, default=40)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--n-target", dest="n_target", type=int, default=8192)
    ap.add_argument("--lanes", type=int, nargs="+", default=[1, 2, 4])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    from autoinst_b200 import api, sharding
    from autoinst_b200.synthetic import CONFIGS, make_chunk, make_map
    dev = torch.device("cuda:0")
    cfg = CONFIGS["tarl_spatial"]
    kw = dict(alpha=cfg["alpha"], theta=cfg["theta"], T=cfg["T"])
    if args.batch:
        chunks = [make_chunk(1000 + i, n_target=args.n_target, features="tarl") for i in range(args.batch)]
    else:
        chunks = make_map(args.map, (3000, 12000), features="tarl", seed=1000)
    sizes = [c.n for c in chunks]
    res = {"chunks": len(chunks), "lanes": {}}
    ref = None
    for L in args.lanes:
        shards = sharding.shard_chunks(sizes, L)          # longest first over the lanes
        packs = [api.PackedChunks([chunks[i].points for i in sh], [chunks[i].tarl for i in sh], None, theta=cfg["theta"], pin=True)
                 for sh in shards]
        devs = [pk.to_device(dev) for pk in packs]

        def run():
            if L == 1:
                api.segment_packed(packs[0], dev_chunks=devs[0], **kw)
            else:
                api.segment_packed_lanes(packs, devs, **kw)
            torch.cuda.synchronize(dev)
        for _ in range(2):
            run()
        ts = []
        for _ in range(args.reps):
            t0 = time.perf_counter()
            run()
            ts.append(1e3 * (time.perf_counter() - t0))
        labs = {}
        for sh, d, pk in zip(shards, devs, packs):
            arr = d.labels.cpu().numpy()
            for k, i in enumerate(sh):
                labs[i] = arr[pk.off[k]:pk.off[k + 1]].copy()
        if ref is None:
            ref = labs
        same = all(np.array_equal(ref[i], labs[i]) for i in ref)
        res["lanes"][L] = {"ms": float(np.median(ts)), "chunks_per_s": len(chunks) / (np.median(ts) / 1e3), "same_labels": bool(same)}
        print(L, res["lanes"][L], flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
