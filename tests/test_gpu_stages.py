"""GPU parity tests, stage by stage, through the C ABI (ctypes) against the CPU oracle."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import CONFIG_NAMES, GOLDEN_SEEDS, load_golden
from autoinst_b200.synthetic import CONFIGS, make_chunk, small_chunk
from oracle import ncut_ref as R
from oracle.affinity_ref import affinity_ref

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _api():
    from autoinst_b200 import api
    return api


def affinity_check(W_gpu, A_ref, tol=1e-5):
    """Parity level 1: identical zero pattern, relative error <= tol on the non-zeros."""
    W = W_gpu.double().cpu().numpy()
    assert np.array_equal(W != 0, A_ref != 0), f"zero pattern differs in {(np.not_equal(W != 0, A_ref != 0)).sum()} entries"
    nz = A_ref != 0
    rel = np.abs(W[nz] - A_ref[nz]) / np.abs(A_ref[nz])
    assert rel.max() <= tol, f"max rel err {rel.max():.3e}"
    return rel.max()


@pytest.mark.parametrize("seed", GOLDEN_SEEDS)
@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_affinity_matches_reference_golden(cuda_device, seed, name):
    inp, out, A = load_golden(seed, name)
    W = _api().affinity(inp["points"], inp["tarl"], inp["dino"], alpha=float(out["alpha"]), theta=float(out["theta"]),
                        gamma=float(out["gamma"]), device=cuda_device)
    affinity_check(W, A.toarray())


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_affinity_matches_oracle_seeded(cuda_device, name):
    cfg = CONFIGS[name]
    ch = make_chunk(5, n_target=2500, features="tarl_dino")       # N not a multiple of the tile
    A = affinity_ref(ch.points, ch.tarl, ch.dino, alpha=cfg["alpha"], theta=cfg["theta"], gamma=cfg["gamma"])
    W, rs = _api().affinity(ch.points, ch.tarl, ch.dino, alpha=cfg["alpha"], theta=cfg["theta"], gamma=cfg["gamma"],
                            device=cuda_device, return_rowsum=True)
    affinity_check(W, A)
    assert np.allclose(rs.cpu().numpy(), A.astype(np.float32).astype(np.float64).sum(1), rtol=1e-6)


def test_affinity_edge_cases(cuda_device):
    api = _api()
    pts = np.array([[0, 0, 0], [1.0, 0, 0], [2.0000001, 0, 0], [0, 0.5, 0]])
    tarl = np.zeros((4, 96), dtype=np.float32); tarl[1] = 1.0; tarl[3] = 1.0
    dino = np.zeros((4, 384), dtype=np.float32); dino[1] = 1.0
    for kw in (dict(alpha=1.0), dict(alpha=0.0), dict(alpha=1.0, theta=0.5), dict(alpha=1.0, theta=0.5, gamma=0.1)):
        A = affinity_ref(pts, tarl.astype(np.float64), dino.astype(np.float64), **kw)
        W = api.affinity(pts, tarl, dino, device=cuda_device, **kw)
        affinity_check(W, A)
    assert float(api.affinity(pts, device=cuda_device)[0, 1]) == pytest.approx(np.exp(-1.0), rel=1e-6)   # inclusive <=
    with pytest.raises(ValueError):
        api.affinity(pts, None, None, gamma=0.1, device=cuda_device)
    one = api.affinity(pts[:1], device=cuda_device)
    assert one.shape == (1, 1) and float(one[0, 0]) == 1.0


def test_degree_and_normalisation(cuda_device):
    api = _api()
    ch = make_chunk(6, n_target=1500, features="tarl")
    A = affinity_ref(ch.points, ch.tarl, alpha=1.0, theta=0.5)
    W = api.affinity(ch.points, ch.tarl, alpha=1.0, theta=0.5, device=cuda_device)
    deg, M = api.degree_normalize(W, return_normalized=True)
    W64 = W.double().cpu().numpy()
    d_ref = 1.0 + W64.sum(0)                                           # normalized_cut.py:38,42
    assert np.allclose(deg.cpu().numpy(), d_ref, rtol=1e-13, atol=0)
    assert np.allclose(deg.cpu().numpy(), 1.0 + A.sum(0), rtol=1e-6)
    M_ref = (W64 + np.eye(ch.n)) / np.sqrt(np.outer(d_ref, d_ref))     # :43-47 (I - M is the Laplacian)
    assert np.allclose(M.double().cpu().numpy(), M_ref, rtol=2e-7, atol=1e-12)


def _blocks(cuda_device, seeds=(31, 32, 33), ppo=450):
    """Block-diagonal matrix of connected single-object blocks, and the per-block oracle results."""
    api = _api()
    mats, offs, ns = [], [], []
    off = 0
    for s in seeds:
        ch = small_chunk(s, n_obj=1, pts_per_obj=ppo + 50 * (s % 3), features="tarl")
        A = affinity_ref(ch.points, ch.tarl, alpha=1.0, theta=0.5)
        ncomp, _ = sp.csgraph.connected_components(sp.csr_matrix(A))
        assert ncomp == 1
        mats.append(A.astype(np.float32))
        offs.append(off); ns.append(ch.n); off += ch.n + (s % 3)       # gaps: unaligned offsets, filler leaves
    n_total = off + 5
    W = np.zeros((n_total, n_total), dtype=np.float32)
    for A, o, n in zip(mats, offs, ns):
        W[o:o + n, o:o + n] = A
    Wd = torch.as_tensor(W, device=cuda_device)
    return api, Wd, W, mats, offs, ns


@pytest.mark.parametrize("impl", [0, 1])        # 0: persistent cluster kernel, 1: grid-wide multi-launch path
def test_lanczos_fiedler_matches_arpack(cuda_device, impl):
    api, Wd, W, mats, offs, ns = _blocks(cuda_device)
    ev, lam2, steps, conv = api.lanczos_fiedler(Wd, offs, ns, lanczos_impl=impl)
    ev = ev.cpu().numpy()
    for A32, o, n, l2, k, c in zip(mats, offs, ns, lam2, steps, conv):
        w = sp.csr_matrix(A32.astype(np.float64))
        with R.pinned_eigsh():
            d, D, ev_ref, vals = R.fiedler_of_block(w)
        assert c == 1 and 0 < k <= n - 1
        assert abs(l2 - vals[1]) < 1e-9, (l2, vals)
        got = ev[o:o + n]
        assert abs(np.linalg.norm(got) - 1) < 1e-12 and got.sum() >= 0
        assert np.abs(got - R.canonical_sign(ev_ref)).max() < 1e-8, np.abs(got - R.canonical_sign(ev_ref)).max()


def test_lanczos_cluster_sizes(cuda_device):
    """One connected block per cluster-size class (1, 2, 4, 8, 16 CTAs) plus one above the cluster limit."""
    api = _api()
    from autoinst_b200.synthetic import make_chunk
    import scipy.sparse.csgraph as csg
    ch = make_chunk(41, n_target=9000, features="tarl")
    A = affinity_ref(ch.points, ch.tarl, alpha=1.0, theta=0.5)
    ncomp, lab = csg.connected_components(sp.csr_matrix(A))
    sizes = np.bincount(lab)
    picks = {}
    for c in np.argsort(-sizes):
        n = int(sizes[c])
        cls = 0 if n <= 320 else 1 if n <= 512 else 2 if n <= 640 else 3 if n <= 1024 else 4 if n <= 2048 else 5
        picks.setdefault(cls, c)
    assert len(picks) >= 3
    mats, offs, ns = [], [], []
    off = 1
    for cls, c in sorted(picks.items()):
        idx = np.where(lab == c)[0]
        mats.append(A[np.ix_(idx, idx)].astype(np.float32)); offs.append(off); ns.append(len(idx)); off += len(idx) + 3
    W = np.zeros((off, off), dtype=np.float32)
    for B, o, n in zip(mats, offs, ns):
        W[o:o + n, o:o + n] = B
    ev, lam2, steps, conv = api.lanczos_fiedler(torch.as_tensor(W, device=cuda_device), offs, ns)
    ev = ev.cpu().numpy()
    for B, o, n, l2, c in zip(mats, offs, ns, lam2, conv):
        with R.pinned_eigsh():
            d, D, ev_ref, vals = R.fiedler_of_block(sp.csr_matrix(B.astype(np.float64)))
        assert c == 1 and abs(l2 - vals[1]) < 1e-9
        assert np.abs(ev[o:o + n] - R.canonical_sign(ev_ref)).max() < 1e-7


@pytest.mark.parametrize("impl", [0, 1])
def test_lanczos_tiny_nodes(cuda_device, impl):
    """n = 3 .. 6: the Krylov space is exhausted after n-1 steps and the result is exact."""
    api = _api()
    rng = np.random.default_rng(0)
    offs, ns, mats = [], [], []
    off = 0
    for n in (3, 4, 5, 6):
        B = rng.uniform(0.1, 0.9, size=(n, n)); B = ((B + B.T) / 2).astype(np.float32); np.fill_diagonal(B, 1.0)
        mats.append(B); offs.append(off); ns.append(n); off += n
    W = np.zeros((off, off), dtype=np.float32)
    for B, o, n in zip(mats, offs, ns):
        W[o:o + n, o:o + n] = B
    ev, lam2, steps, conv = api.lanczos_fiedler(torch.as_tensor(W, device=cuda_device), offs, ns, lanczos_impl=impl)
    ev = ev.cpu().numpy()
    for B, o, n, l2 in zip(mats, offs, ns, lam2):
        Wf = B.astype(np.float64) + np.eye(n)
        d = Wf.sum(0)
        L = np.eye(n) - Wf / np.sqrt(np.outer(d, d))
        vals, vecs = np.linalg.eigh(L)
        assert abs(l2 - vals[1]) < 1e-12
        assert np.abs(ev[o:o + n] - R.canonical_sign(vecs[:, 1])).max() < 1e-9


def test_ncut_scan_matches_reference_costs(cuda_device):
    api, Wd, W, mats, offs, ns = _blocks(cuda_device)
    ev_full = np.zeros(W.shape[0])
    refs = []
    for A32, o, n in zip(mats, offs, ns):
        w = sp.csr_matrix(A32.astype(np.float64))
        with R.pinned_eigsh():
            d, D, ev, vals = R.fiedler_of_block(w)
        ev = R.canonical_sign(ev)
        ev_full[o:o + n] = ev
        costs = [R._ncut_value(w, D, d, ev > t, False) for t in np.linspace(ev.min(), ev.max(), 10, endpoint=False)]
        side, best = R.best_threshold_cut(ev, D, d, w)
        refs.append((np.array(costs), side, best))
    best_k, mcut, costs, mask = api.ncut_scan(Wd, offs, ns, ev_full)
    mask = mask.cpu().numpy().astype(bool)
    for (c_ref, side, best), o, n, bk, mc, cs in zip(refs, offs, ns, best_k, mcut, costs):
        assert np.allclose(cs, c_ref, rtol=1e-9), (cs, c_ref)
        assert bk == int(np.argmin(c_ref)) and abs(mc - best) <= 1e-9 * best
        assert np.array_equal(mask[o:o + n], side)


def test_ncut_scan_allclose_and_ties(cuda_device):
    api = _api()
    n = 6
    W = torch.ones((n, n), device=cuda_device)
    best_k, mcut, costs, mask = api.ncut_scan(W, [0], [n], np.full(n, 1 / np.sqrt(n)))
    assert best_k[0] == -1 and np.isinf(mcut[0])                       # np.allclose(mn, mx): normalized_cut.py:22-23
    ev = np.array([-0.5, -0.5, -0.5, 0.5, 0.5, 0.5])                   # thresholds 1..9 give one mask: first wins (:30)
    best_k, mcut, costs, mask = api.ncut_scan(W, [0], [n], ev)
    w = sp.csr_matrix(np.ones((n, n)))
    d = np.full(n, n + 1.0)
    side, best = R.best_threshold_cut(ev, sp.diags(d), d, w)
    assert best_k[0] == 0 and abs(mcut[0] - best) < 1e-12
    assert np.array_equal(mask.cpu().numpy().astype(bool), side)
    assert np.all(costs[0] == costs[0][0])                             # identical masks -> bit-identical costs


def test_partition_matches_fancy_indexing(cuda_device):
    api, Wd, W, mats, offs, ns = _blocks(cuda_device)
    rng = np.random.default_rng(5)
    mask = np.zeros(W.shape[0], dtype=np.uint8)
    for o, n in zip(offs, ns):
        mask[o:o + n] = rng.random(n) < 0.4
    # without component splitting: exactly w[mask][:, mask] and w[~mask][:, ~mask] (normalized_cut.py:57-58)
    Wo, perm, coff, cn = api.partition(Wd, offs, ns, mask, split_components=False)
    Wo = Wo.cpu().numpy(); perm = perm.cpu().numpy()
    assert sorted(perm.tolist()) == list(range(W.shape[0]))
    assert len(coff) == 2 * len(offs)
    ci = 0
    for o, n in zip(offs, ns):
        m = mask[o:o + n].astype(bool)
        for sel in (m, ~m):
            idx = o + np.where(sel)[0]
            co, cnn = coff[ci], cn[ci]; ci += 1
            assert cnn == len(idx) and np.array_equal(perm[co:co + cnn], idx)
            assert np.array_equal(Wo[co:co + cnn, co:co + cnn], W[np.ix_(idx, idx)])
    # with component splitting: children are the connected components of each side
    Wo, perm, coff, cn = api.partition(Wd, offs, ns, mask, split_components=True)
    Wo = Wo.cpu().numpy(); perm = perm.cpu().numpy()
    for co, cnn in zip(coff, cn):
        idx = perm[co:co + cnn]
        if cnn > 2:
            assert np.array_equal(Wo[co:co + cnn, co:co + cnn], W[np.ix_(idx, idx)])
        ncomp, _ = sp.csgraph.connected_components(sp.csr_matrix(W[np.ix_(idx, idx)] != 0))
        assert ncomp == 1
        assert len(set(mask[idx].tolist())) == 1
    total = sum(cn)
    assert total == sum(ns)
    expected = 0
    for o, n in zip(offs, ns):
        m = mask[o:o + n].astype(bool)
        for sel in (m, ~m):
            idx = o + np.where(sel)[0]
            expected += sp.csgraph.connected_components(sp.csr_matrix(W[np.ix_(idx, idx)] != 0))[0]
    assert len(coff) == expected


@pytest.mark.parametrize("name", ["tarl_spatial", "tarl_spatial_dino"])
def test_affinity_tensor_core_path(cuda_device, name):
    """affinity_impl=1: TARL Gram matrix by tcgen05 (3xTF32, TMA-fed) must meet the same level-1 gate."""
    cfg = CONFIGS[name]
    api = _api()
    for seed in GOLDEN_SEEDS:
        inp, out, A = load_golden(seed, name)
        W = api.affinity(inp["points"], inp["tarl"], inp["dino"], alpha=float(out["alpha"]), theta=float(out["theta"]),
                         gamma=float(out["gamma"]), device=cuda_device, impl=1)
        affinity_check(W, A.toarray())
    ch = make_chunk(5, n_target=2500, features="tarl_dino")
    A = affinity_ref(ch.points, ch.tarl, ch.dino, alpha=cfg["alpha"], theta=cfg["theta"], gamma=cfg["gamma"])
    W1 = api.affinity(ch.points, ch.tarl, ch.dino, alpha=cfg["alpha"], theta=cfg["theta"], gamma=cfg["gamma"],
                      device=cuda_device, impl=1)
    affinity_check(W1, A)
    # identical feature rows (distance 0) and near-identical rows: the cancellation fallback must hold the gate
    t2 = ch.tarl.copy()
    t2[1] = t2[0]
    t2[3] = t2[2] * (1 + 1e-6)
    A2 = affinity_ref(ch.points, t2, None, alpha=1.0, theta=0.5)
    W2 = api.affinity(ch.points, t2, None, alpha=1.0, theta=0.5, device=cuda_device, impl=1)
    affinity_check(W2, A2)
    # no feature term: nothing for the tensor cores to do, the exact tile kernel runs
    affinity_check(api.affinity(ch.points, None, None, alpha=1.0, device=cuda_device, impl=1), affinity_ref(ch.points, alpha=1.0))
    with pytest.raises(Exception):                                      # DINOv2 term without TARL: fails loudly
        api.affinity(ch.points, None, ch.dino, alpha=1.0, gamma=0.1, device=cuda_device, impl=1)


def test_nn_reprojection_matches_kdtree(cuda_device):
    """next-row N1: 1-NN label re-projection (point_cloud_utils.py:144-174) against scipy's KD-tree."""
    from scipy.spatial import cKDTree
    api = _api()
    rng = np.random.default_rng(3)
    src = rng.uniform(-12.5, 12.5, size=(3000, 3))
    lab = rng.integers(0, 40, size=3000).astype(np.int32)
    qry = src[rng.integers(0, 3000, size=20000)] + rng.normal(0, 0.2, size=(20000, 3))
    dist, idx = cKDTree(src).query(qry, k=1)
    out, gi = api.nn_reproject(qry, src, lab, device=cuda_device)
    assert np.array_equal(gi.cpu().numpy(), idx)
    assert np.array_equal(out.cpu().numpy(), lab[idx])
    out_r, _ = api.nn_reproject(qry, src, lab, max_radius=0.3, no_label=-1, device=cuda_device)
    expect = np.where(dist > 0.3, -1, lab[idx])
    assert np.array_equal(out_r.cpu().numpy(), expect)
