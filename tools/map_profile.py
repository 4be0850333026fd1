"""Where the map-level post-processing (N2 merge, remove_semantics, N4 metrics) of one synthetic map spends its time:
host timer with a device synchronisation around each part, `--reps` repetitions after one warm-up.
    python tools/map_profile.py [--chunks 40] [--out gpurun_out/map_profile.json]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=40)
    ap.add_argument("--lo", type=int, default=3000)
    ap.add_argument("--hi", type=int, default=12000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    from autoinst_b200 import api
    from autoinst_b200.synthetic import CONFIGS, make_map
    dev = torch.device("cuda:0")
    cfg = CONFIGS["tarl_spatial"]
    chunks = make_map(args.chunks, (args.lo, args.hi), features="tarl", seed=1000)
    packed = api.PackedChunks([c.points for c in chunks], [c.tarl for c in chunks], None, theta=cfg["theta"], pin=True)
    dc = packed.to_device(dev)
    api.segment_packed(packed, dev_chunks=dc, alpha=cfg["alpha"], theta=cfg["theta"], T=cfg["T"])
    seg = dc.labels
    post = api.MapPost(chunks, device=dev)
    hd = api.Handle.get(dev)

    def timed(fn):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize(dev)
        return out, 1e3 * (time.perf_counter() - t0)

    rows = {"global_labels": [], "merge": [], "remove_semantics": [], "metrics": []}
    launches = {}
    for rep in range(args.reps + 1):
        hd.launch_count(reset=True)
        lab, t_a = timed(lambda: post.global_labels(seg))
        launches["global_labels"] = hd.launch_count(reset=True)
        (out_lab, idx, kept), t_b = timed(lambda: api._merge_device(hd, post.off, post.pts, lab, post.centers, post.half,
                                                                    post.min_iou, dev))
        launches["merge"] = hd.launch_count(reset=True)
        keep = idx[:kept]
        merged = out_lab[keep].contiguous()
        gt = post.gt_all[keep].contiguous()
        inst = torch.empty_like(merged)

        def rs():
            with torch.cuda.device(dev):
                api.check(hd.lib.ancuts_remove_semantics(hd.h, int(merged.numel()), api._ptr(gt), api._ptr(merged),
                                                         post.threshold, api._ptr(inst), api._stream(dev)))
        _, t_c = timed(rs)
        launches["remove_semantics"] = hd.launch_count(reset=True)
        m, t_d = timed(lambda: api.instance_metrics(merged, inst, gt, min_points=post.min_points, device=dev))
        launches["metrics"] = hd.launch_count(reset=True)
        if rep:
            for k, t in zip(rows, (t_a, t_b, t_c, t_d)):
                rows[k].append(t)
    res = {"chunks": args.chunks, "points": int(sum(c.n for c in chunks)), "kept": int(kept),
           "ms": {k: float(np.mean(v)) for k, v in rows.items()}, "launches": {k: int(v) for k, v in launches.items()},
           "metrics": {k: float(v) for k, v in m.items()}}
    print(json.dumps(res))
    if args.out:
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
