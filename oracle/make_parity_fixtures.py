#!/usr/bin/env python
"""Oracle labels at the BENCHMARKED size for the driver-run `-m gpu` parity tests (TEST INFRASTRUCTURE).

    python -m oracle.make_parity_fixtures            # writes tests/golden/parity_bench_size.npz

For every (config, n_target, seed, clutter) below: `synthetic.make_chunk` -> `oracle.affinity_ref` ->
`oracle.ncut_ref.normalized_cut_ref` under the eigsh pin (`v0 = ones`), and once more with a random start
vector to mark the chunk oracle-stable or not.  The oracle needs 25-90 s per chunk, far
too long for the GPU suite, so the labels (int16) travel as a fixture; the inputs are regenerated from the seed.
`tools/parity_sweep.py --oracle-only --oracle-cache parity_cache` fills the same cache in parallel.
The restatements themselves are pinned against the unmodified reference by `oracle/make_golden.py`.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# (config, n_target, first seed, count, clutter)
SETS = [("spatial", 8192, 7000, 6, 0), ("tarl_spatial", 8192, 7000, 6, 0), ("tarl_spatial_dino", 8192, 7000, 6, 0),
        ("tarl_spatial", 16384, 7200, 3, 0), ("tarl_spatial", 8192, 7400, 8, 40)]
OUT = os.path.join(ROOT, "tests", "golden", "parity_bench_size.npz")


def tag(name, n_target, seed, clutter):
    return f"{name}_{n_target}_{seed}" + (f"_c{clutter}" if clutter else "")


def main():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from parity_sweep import oracle_job
    cache = os.path.join(ROOT, "parity_cache")
    out = {}
    for name, n_target, seed0, count, clutter in SETS:
        for s in range(seed0, seed0 + count):
            _, labels, stable = oracle_job((s, n_target, name, cache, clutter))
            assert labels.max() < 32767
            t = tag(name, n_target, s, clutter)
            out["labels_" + t] = labels.astype(np.int16)
            out["stable_" + t] = np.bool_(stable)
            print(t, labels.shape[0], int(labels.max()) + 1, bool(stable), flush=True)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
