"""Digest of what one segment call returns for the bench batch (labels, node records): to compare two builds of the
library (ANCUTS_LIB_PATH) that are meant to give bit-identical results.
    python tools/labels_digest.py [--batch 32]"""
import argparse
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--n-target", dest="n_target", type=int, default=8192)
    args = ap.parse_args()
    import torch
    from autoinst_b200 import api
    from autoinst_b200.synthetic import CONFIGS, make_chunk
    cfg = CONFIGS["tarl_spatial"]
    chunks = [make_chunk(1000 + i, n_target=args.n_target, features="tarl") for i in range(args.batch)]
    packed = api.PackedChunks([c.points for c in chunks], [c.tarl for c in chunks], None, theta=cfg["theta"])
    res = api.segment_packed(packed, device=torch.device("cuda:0"), alpha=cfg["alpha"], theta=cfg["theta"], T=cfg["T"],
                             want_stats=True)
    h = hashlib.sha1()
    for lab in res.labels:
        h.update(np.ascontiguousarray(lab).tobytes())
    st = res.stats
    order = np.lexsort((st["n_side"], st["n"], st["level"], st["chunk"]))
    h2 = hashlib.sha1()
    for k in ("chunk", "level", "n", "n_side", "best_k", "split", "steps"):
        h2.update(np.ascontiguousarray(st[k][order]).tobytes())
    h3 = hashlib.sha1(np.ascontiguousarray(st["mcut"][order]).tobytes())
    print("labels", h.hexdigest()[:16], "records", h2.hexdigest()[:16], "mcut", h3.hexdigest()[:16], "nodes", len(st),
          "steps", int(st["steps"].sum()), "unconverged", res.unconverged)


if __name__ == "__main__":
    main()
