"""GPU tests of the kernel variants behind the same C ABI: two-pass affinity and its overflow fallback, the
Lanczos matvec / Gram-Schmidt variants (ANCUTS_X, read when a handle is created), the per-level trace."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from autoinst_b200.synthetic import CONFIGS, make_chunk
from oracle import ncut_ref as R
from oracle.affinity_ref import affinity_ref

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _api():
    from autoinst_b200 import api
    return api


def variant_lane(dev, flags, lane):
    """A library handle created with ANCUTS_X=flags (its own workspace), addressed by `lane`."""
    api = _api()
    old = os.environ.get("ANCUTS_X")
    os.environ["ANCUTS_X"] = str(flags)
    try:
        api.Handle.get(dev, lane)
    finally:
        if old is None:
            del os.environ["ANCUTS_X"]
        else:
            os.environ["ANCUTS_X"] = old
    return lane


def check_affinity(W, A):
    W = W.double().cpu().numpy()
    assert np.array_equal(W != 0, A != 0)
    rel = np.abs(W - A)[A != 0] / A[A != 0]
    assert rel.max() <= 1e-5, rel.max()              # north_star level 1
    assert np.array_equal(W, W.T)


# ANCUTS_X bits (engine.cu): 2 = integer widening of every second element, 32 = L2 prefetch two passes ahead, 256 = one-kernel
# affinity, 1024 = three-term + one Gram-Schmidt pass, 4096 = basis rows in global memory only, 8192 = TMA ring,
# 16384 = adaptive convergence checks, 32768 = division-free Sturm counts, 65536 = 128 shifts per round,
# 131072 = start vector from the coordinates, 262144 = deferred affinity (W written block by block after the root split),
# 524288 = pairs of the deferred affinity from a cell grid instead of the tile sweep.
S3 = 2 | 1024 | 8192 | 16384 | 32768 | 65536 | 131072
VARIANTS = {"session3_default": S3 | 262144, "session3_grid_pairs": S3 | 262144 | 524288, "session3_dense_affinity": S3, "session3_prefetch_next_matvec": S3 | 262144 | 16,
            "session3_hash_start_256_shifts": 2 | 1024 | 8192 | 16384 | 32768 | 262144,
            "default": 2 | 1024 | 8192, "one_kernel_affinity": 2 | 1024 | 8192 | 256, "register_matvec_prefetch": 2 | 32 | 1024,
            "register_matvec_cgs2": 0, "ring_cgs2": 8192, "basis_in_global": 2 | 1024 | 8192 | 4096}


@pytest.mark.parametrize("variant", list(VARIANTS))
def test_variants_give_oracle_labels(cuda_device, variant):
    api = _api()
    lane = variant_lane(cuda_device, VARIANTS[variant], 10 + list(VARIANTS).index(variant))
    cfg = CONFIGS["tarl_spatial"]
    chunks = [make_chunk(300 + i, n_target=1500 + 700 * i, features="tarl") for i in range(3)]
    packed = api.PackedChunks([c.points for c in chunks], [c.tarl for c in chunks], None, theta=cfg["theta"])
    res = api.segment_packed(packed, device=cuda_device, lane=lane, alpha=cfg["alpha"], theta=cfg["theta"], T=cfg["T"],
                             want_stats=True)
    assert int((res.stats["converged"] == 0).sum()) == 0
    for ch, lab in zip(chunks, res.labels):
        A = affinity_ref(ch.points, ch.tarl, None, alpha=cfg["alpha"], theta=cfg["theta"])
        with R.pinned_eigsh():
            g = R.normalized_cut_ref(sp.csr_matrix(A), ch.n, np.arange(ch.n), T=cfg["T"])
        assert R.same_partition(lab, R.labels_from_groups(g, ch.n)), variant


@pytest.mark.parametrize("name", ["tarl_spatial", "tarl_spatial_dino"])
def test_two_pass_and_one_kernel_affinity_agree_with_the_oracle(cuda_device, name):
    api = _api()
    cfg = CONFIGS[name]
    ch = make_chunk(41, n_target=2500, features="tarl_dino" if cfg["gamma"] else "tarl")
    A = affinity_ref(ch.points, ch.tarl, ch.dino if cfg["gamma"] else None, alpha=cfg["alpha"], theta=cfg["theta"], gamma=cfg["gamma"])
    lane1 = variant_lane(cuda_device, VARIANTS["one_kernel_affinity"], 10 + list(VARIANTS).index("one_kernel_affinity"))
    for lane in (0, lane1):
        W = api.affinity(ch.points, ch.tarl, ch.dino if cfg["gamma"] else None, alpha=cfg["alpha"], theta=cfg["theta"],
                         gamma=cfg["gamma"], device=cuda_device, lane=lane)
        check_affinity(W, A)


def test_pair_queue_overflow_falls_back(cuda_device):
    """More than 96 in-mask pairs per point on average: the queue of the two-pass form overflows and the one-kernel
    form takes over, in the stage entry point (conditional launch) and in the segment call (host read-back)."""
    api = _api()
    rng = np.random.default_rng(5)
    n = 900
    pts = (rng.normal(size=(n, 3)) * 0.2).astype(np.float32).astype(np.float64)       # one 1 m blob: almost every pair inside
    pts[n // 2:] += 7.0                                                                  # ... and a second one far away
    tarl = rng.normal(size=(n, 96)).astype(np.float32)
    A = affinity_ref(pts, tarl.astype(np.float64), None, alpha=1.0, theta=0.5)
    assert (A != 0).sum() > 2 * 96 * n
    W = api.affinity(pts, tarl, alpha=1.0, theta=0.5, device=cuda_device)
    check_affinity(W, A)
    lab = api.segment_chunk(pts, tarl, alpha=1.0, theta=0.5, T=0.03, device=cuda_device)
    with R.pinned_eigsh():
        g = R.normalized_cut_ref(sp.csr_matrix(A), n, np.arange(n), T=0.03)
    assert R.same_partition(lab, R.labels_from_groups(g, n))


def test_level_trace(cuda_device):
    api = _api()
    cfg = CONFIGS["tarl_spatial"]
    chunks = [make_chunk(500 + i, n_target=2000, features="tarl") for i in range(2)]
    packed = api.PackedChunks([c.points for c in chunks], [c.tarl for c in chunks], None, theta=cfg["theta"])
    devc = packed.to_device(cuda_device)
    hd = api.Handle.get(cuda_device)
    hd.set_stage_timing(2)
    try:
        res = api.segment_packed(packed, dev_chunks=devc, alpha=cfg["alpha"], theta=cfg["theta"], T=cfg["T"], want_stats=True)
        lv = hd.levels()
        acc = hd.accounting()
    finally:
        hd.set_stage_timing(0)
    assert len(lv) >= 1 and all(l["ms"] > 0 for l in lv)
    assert sum(sum(l["bins"]) + l["big"] for l in lv) == len(res.stats)
    assert abs(sum(l["ms"] for l in lv) - acc["matvec"]["ms"]) < 1e-3 * max(acc["matvec"]["ms"], 1.0) + 1e-3
    n = res.stats["n"].astype(np.float64)
    k = res.stats["steps"].astype(np.float64)
    assert acc["matvec"]["bytes"] == pytest.approx(float((k * (4 * n * n + 8 * n)).sum()), rel=1e-12)
