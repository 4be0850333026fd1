"""CPU tests: the oracle against the golden vectors generated from the unmodified reference
(oracle/make_golden.py), and the numpy model of the device algorithm against the oracle."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import CONFIG_NAMES, GOLDEN, GOLDEN_SEEDS, load_golden
from autoinst_b200.synthetic import CONFIGS, make_chunk, small_chunk
from oracle import ncut_ref as R
from oracle.affinity_ref import affinity_ref, drop_isolated
from oracle.device_model import lanczos_fiedler, scan_cuts, segment_model


@pytest.mark.parametrize("seed", GOLDEN_SEEDS)
@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_affinity_ref_matches_reference_bitwise(seed, name):
    inp, out, A = load_golden(seed, name)
    got = affinity_ref(inp["points"], inp["tarl"].astype(np.float64), inp["dino"].astype(np.float64),
                       alpha=float(out["alpha"]), theta=float(out["theta"]), gamma=float(out["gamma"]))
    assert np.array_equal(got, A.toarray())
    keep, sub = drop_isolated(got)
    assert len(keep) == got.shape[0]          # A_ii = 1: nothing is ever isolated (point_cloud_utils.py:189-195)


@pytest.mark.parametrize("seed", GOLDEN_SEEDS)
@pytest.mark.parametrize("name", CONFIG_NAMES)
@pytest.mark.parametrize("faithful", [True, False])
def test_normalized_cut_ref_matches_reference(seed, name, faithful):
    inp, out, A = load_golden(seed, name)
    n = A.shape[0]
    with R.pinned_eigsh():
        groups = R.normalized_cut_ref(A, n, np.arange(n), T=float(out["T"]), split_lim=0.01, faithful=faithful)
    assert np.array_equal([len(g) for g in groups], out["group_sizes"])     # same DFS order as the reference
    assert np.array_equal(R.labels_from_groups(groups, n), out["labels"])


def test_known_answers():
    kat = np.load(f"{GOLDEN}/known_answers.npz")
    names = sorted({k[:-2] for k in kat.files if k.endswith("_w")})
    assert names
    for nm in names:
        w, lab = kat[nm + "_w"], kat[nm + "_labels"]
        T = float(nm.split("_T")[1])
        with R.pinned_eigsh():
            g = R.normalized_cut_ref(sp.csr_matrix(w), w.shape[0], np.arange(w.shape[0]), T=T)
        assert np.array_equal(R.labels_from_groups(g, w.shape[0]), lab)
    # two cliques joined by a weak edge: hand-computed N-cut of the separating cut (normalized_cut.py:4-11)
    w = kat["two_cliques_T0.03_w"]
    d = 1.0 + w.sum(0)
    side = np.arange(12) >= 6
    cut = w[np.ix_(side, ~side)].sum()
    expect = cut / d[side].sum() + cut / d[~side].sum()
    assert abs(cut - 0.01) < 1e-15
    D = sp.diags(d)
    ev = np.where(side, 1.0, -1.0)
    mask, cost = R.best_threshold_cut(ev, D, d, sp.csr_matrix(w))
    assert np.array_equal(mask, side) and abs(cost - expect) < 1e-15


def test_stop_rules():
    # n <= 2 and n/N <= split_lim are leaves (normalized_cut.py:39-40)
    w = sp.csr_matrix(np.array([[1.0, 0.0], [0.0, 1.0]]))
    assert len(R.normalized_cut_ref(w, 2, np.arange(2), T=1.0)) == 1
    w3 = sp.csr_matrix(np.eye(3))
    assert len(R.normalized_cut_ref(w3, 1000, np.arange(3), T=1.0)) == 1


def test_inclusive_threshold_and_zero_rows():
    pts = np.array([[0, 0, 0], [1.0, 0, 0], [2.0000001, 0, 0]])
    A = affinity_ref(pts, alpha=1.0)
    assert A[0, 1] == np.exp(-1.0) and A[1, 2] == 0.0             # <= 1.0 inclusive, ncuts_utils.py:61
    tarl = np.zeros((3, 96)); tarl[1] = 1.0
    B = affinity_ref(pts, tarl, alpha=1.0, theta=0.5)
    assert B[0, 1] == np.exp(-1.0)                                 # zero row neutralised, :145-146
    dino = np.zeros((3, 384)); dino[1] = 1.0
    Cm = affinity_ref(pts, None, dino, alpha=1.0, gamma=0.1)
    assert np.isclose(Cm[0, 1], np.exp(-1.0 - 0.1 * np.sqrt(384)))  # DINO zero rows are not, :129-133
    with pytest.raises(ValueError):
        affinity_ref(pts, None, None, gamma=0.1)


@pytest.mark.parametrize("seed", GOLDEN_SEEDS)
@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_device_model_matches_reference_on_golden(seed, name):
    inp, out, A = load_golden(seed, name)
    lab = segment_model(A.toarray().astype(np.float32), float(out["T"]))
    assert R.same_partition(lab, out["labels"])


def test_device_model_stage_parity():
    """Lanczos vs ARPACK shift-invert, bucket scan vs ten ncut_cost calls, on one connected block."""
    ch = small_chunk(21, n_obj=1, pts_per_obj=500, features="tarl")
    A = affinity_ref(ch.points, ch.tarl, alpha=1.0, theta=0.5)
    w = sp.csr_matrix(A)
    with R.pinned_eigsh():
        d, D, ev_ref, vals = R.fiedler_of_block(w)
    ev, lam2 = lanczos_fiedler(A.astype(np.float32), 1.0 + A.astype(np.float32).astype(np.float64).sum(1))
    assert abs(lam2 - vals[1]) < 1e-6
    assert np.abs(ev - R.canonical_sign(ev_ref)).max() < 1e-5      # float32 W vs float64 W
    side_ref, cost_ref = R.best_threshold_cut(ev, D, d, w)
    k, cost, bucket = scan_cuts(A.astype(np.float32), d, ev)
    assert np.array_equal(bucket > k, side_ref) and abs(cost - cost_ref) < 1e-6


def test_device_model_matches_oracle_on_a_chunk():
    ch = make_chunk(3, n_target=1500, features="tarl")
    cfg = CONFIGS["tarl_spatial"]
    A = affinity_ref(ch.points, ch.tarl, alpha=cfg["alpha"], theta=cfg["theta"])
    with R.pinned_eigsh():
        g = R.normalized_cut_ref(sp.csr_matrix(A), ch.n, np.arange(ch.n), T=cfg["T"])
    lab = segment_model(A.astype(np.float32), cfg["T"])
    assert R.same_partition(lab, R.labels_from_groups(g, ch.n))


def test_metrics_ref_matches_reference_golden():
    """oracle.metrics_ref against the outputs of the reference's own Metrics class (tests/golden/metrics.npz)."""
    from oracle.metrics_ref import instance_metrics
    g = np.load(f"{GOLDEN}/metrics.npz")
    keys = ("p", "r", "f1", "ap", "ap0.25", "ap0.5", "S_assoc")
    for case in range(4):
        for mp in (200, 20):
            got = instance_metrics(g[f"c{case}_all"], g[f"c{case}_pred"], g[f"c{case}_gt"], min_points=mp)
            assert np.allclose([got[k] for k in keys], g[f"c{case}_mp{mp}_ref"], rtol=0, atol=1e-12)


def test_merge_ref_unites_instances_across_overlapping_chunks():
    from autoinst_b200.synthetic import make_map
    from oracle import merge_ref as M
    chunks = make_map(4, 1500, seed=5)
    # GT labels as "predictions": every instance seen by two chunks must end up with ONE merged label
    parts = [(c.points, M.globally_unique(c.chunk_id, c.instance) * (c.instance != 0)) for c in chunks]
    pts, lab = M.merge_chunks_unite_instances(parts)
    gpts, glab = M.merge_unite_gt([(c.points, c.instance) for c in chunks])
    assert pts.shape == gpts.shape and np.array_equal(pts, gpts)          # same de-duplicated map
    assert len(pts) < sum(c.n for c in chunks)                            # the 3 m overlaps were de-duplicated
    for g in np.unique(glab):
        if g == 0:
            continue
        assert len(np.unique(lab[glab == g])) == 1, g
    pred = M.compact_labels(lab)
    assert pred.min() == 0 and (pred == 0).sum() == (glab == 0).sum()
    cleaned = M.remove_semantics(M.compact_labels(glab), pred)
    assert np.array_equal(cleaned, pred)                                   # nothing sits on GT background here


@pytest.mark.parametrize("seed,expect_groups", [(42, 0), (43, 1)])
def test_disconnected_nodes_peel_one_component_and_leave_a_residual_group(seed, expect_groups):
    """What the reference does with tiny fragments: every split of a disconnected node peels ONE
    connected component (the null vectors of the block-diagonal Laplacian are localised on components and
    `eigsh` returns the two with the largest rounding residue, `normalized_cut.py:49-53`), and the chain ends in
    one leaf of whole components once it is at most 1 % of N (`:39-40`).  The device algorithm gives every
    component its own segment, so it equals the reference up to that residual leaf."""
    from scipy.sparse.csgraph import connected_components
    cfg = CONFIGS["tarl_spatial"]
    ch = make_chunk(seed, n_target=1500, features="tarl", clutter=10)
    A = affinity_ref(ch.points, ch.tarl, alpha=cfg["alpha"], theta=cfg["theta"])
    w = sp.csr_matrix(A)
    ncomp, cc = connected_components(w, directed=False)
    sizes = set(np.bincount(cc).tolist())
    trace = []
    with R.pinned_eigsh():
        g = R.normalized_cut_ref(w, ch.n, np.arange(ch.n), T=cfg["T"], trace=trace)
    ref = R.labels_from_groups(g, ch.n)
    degenerate = [t for t in trace if "vals" in t and abs(t["vals"][1]) < 1e-12]
    assert len(degenerate) >= ncomp - 3
    for t in degenerate:
        assert t["split"] and abs(t["mcut"]) < 1e-9
        assert t["n_side"] in sizes or (t["n"] - t["n_side"]) in sizes     # one whole component leaves the node
    lab = segment_model(A.astype(np.float32), cfg["T"])
    rep = R.residual_group_report(lab, ref, ch.n)
    assert rep["refines"] and rep["residual_only"]
    assert len(rep["groups"]) == expect_groups
    for _, _, pts in rep["groups"]:
        assert pts <= 0.01 * ch.n


def test_merge_ref_matches_the_reference_golden():
    """tests/golden/merge.npz: outputs of the reference's own merge_chunks_unite_instances2 / merge_unite_gt /
    remove_semantics (`point_cloud_utils.py:253-287,320-329,387-491`), run unmodified by oracle/make_golden.py with a
    numpy stand-in for the Open3D PointCloud (crop, +=, remove_duplicated_points)."""
    from oracle import merge_ref as M
    g = np.load(f"{GOLDEN}/merge.npz")
    n = int(g["n_chunks"])
    parts = [(g[f"p{i}"], g[f"l{i}"]) for i in range(n)]
    pts, lab = M.merge_chunks_unite_instances(parts)
    assert np.array_equal(pts, g["merged_points"]) and np.array_equal(lab, g["merged_labels"])
    assert len(np.unique(lab)) < sum(len(np.unique(l)) for _, l in parts) - n          # instances were united across chunks
    gpts, glab = M.merge_unite_gt([(g[f"p{i}"], g[f"g{i}"]) for i in range(n)])
    assert np.array_equal(gpts, pts) and np.array_equal(glab, g["gt_labels"])
    cleaned = M.remove_semantics(M.compact_labels(glab), M.compact_labels(lab))
    assert np.array_equal(cleaned, g["cleaned"])
