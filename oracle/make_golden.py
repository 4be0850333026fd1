"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

    python -m oracle.make_golden            # needs /root/reference, never runs on the GPU box

What it pins (SURVEY.md §8c — the reference has no tests of its own for this path):

1. affinity: the reference's `ncuts_chunk` (`pipeline/ncuts/ncuts_utils.py:28-174`) is imported with
   stub `open3d` / `matplotlib` modules and driven up to its `normalized_cut(...)` call, which is
   intercepted to capture the affinity matrix the reference actually builds.  The feature fetchers
   outside the hot path (`tarl_features_per_patch`, `image_based_features_per_patch`) are replaced
   by the synthetic arrays; `dinov2_mean` is the reference's own.  `oracle.affinity_ref` must
   reproduce that matrix bit for bit, else this script fails.
2. recursion: the reference's `normalized_cut` (`pipeline/ncuts/normalized_cut.py:37-63`) is run on
   that matrix with the eigsh pin of `oracle.ncut_ref.pinned_eigsh`; `oracle.ncut_ref` must return
   the same groups in the same order, else this script fails.
3. the captured matrices and groups are written as fixtures for the CPU and GPU test suites.
4. feature pooling (row N3): the reference's own `tarl_features_per_patch`
   (`pipeline/utils/point_cloud/chunk_generation.py:205-258`) is run unmodified on synthetic scans with a
   stand-in for the three Open3D pieces it touches (`PointCloud.transform`, `Vector3dVector`, `KDTreeFlann.
   search_radius_vector_3d`; the KD-tree stand-in is scipy's cKDTree with nanoflann's published rule: squared
   distance strictly below radius^2, results sorted by distance).  `oracle.pool_ref` must reproduce its output bit
   for bit, else this script fails; inputs and output go to tests/golden/pooling.npz.  What stays unpinned is only
   Open3D's own behaviour at the sphere boundary and its result order (which moves a mean by rounding).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import scipy.sparse as sp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/pipeline"
OUT = os.path.join(REPO, "tests", "golden")


class _Captured(Exception):
    pass


class _FakeCloud:
    def __init__(self, pts):
        self.points = np.asarray(pts, dtype=np.float64)

    def select_by_index(self, idx):
        return _FakeCloud(self.points[np.asarray(idx)])


def import_reference():
    """Import ncuts.ncuts_utils and ncuts.normalized_cut from the reference tree."""
    if not os.path.isdir(REF):
        raise SystemExit("reference tree not found: golden vectors can only be generated in the build container")
    for name in ["open3d", "open3d.geometry", "open3d.utility", "open3d.io", "open3d.pipelines",
                 "open3d.pipelines.registration", "matplotlib", "matplotlib.pyplot"]:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["open3d"].geometry = sys.modules["open3d.geometry"]
    sys.modules["open3d"].utility = sys.modules["open3d.utility"]
    sys.modules["open3d"].io = sys.modules["open3d.io"]
    sys.modules["open3d"].pipelines = sys.modules["open3d.pipelines"]
    sys.modules["open3d.pipelines"].registration = sys.modules["open3d.pipelines.registration"]
    plt = sys.modules["matplotlib.pyplot"]
    plt.cm = types.SimpleNamespace(viridis=lambda x: np.zeros((len(x), 4)))
    sys.modules["matplotlib"].pyplot = plt
    cwd = os.getcwd()
    os.chdir(REF)                       # config.py:79 opens "utils/semantic-kitti.yaml" relative to cwd
    mine = [k for k in sys.modules if k == "ncuts" or k.startswith("ncuts.") or k == "config"]
    stash = {m: sys.modules.pop(m) for m in mine}
    # the reference's ncuts/ has no __init__.py; this repo's own `ncuts` package would shadow it, so the
    # name is bound to the reference directory explicitly while its modules are imported
    pkg = types.ModuleType("ncuts")
    pkg.__path__ = [os.path.join(REF, "ncuts")]
    sys.modules["ncuts"] = pkg
    sys.path.insert(0, REF)
    try:
        import ncuts.ncuts_utils as nu
        import ncuts.normalized_cut as nc
        import config as cfg
        assert nu.__file__.startswith(REF) and nc.__file__.startswith(REF)
    finally:
        os.chdir(cwd)
        sys.path.remove(REF)
    ref = dict(nu=nu, nc=nc, cfg=cfg)
    # leave the reference modules out of sys.modules so the repo's own `ncuts` package stays importable
    for m in [k for k in sys.modules if k == "ncuts" or k.startswith("ncuts.") or k == "config"
              or k == "utils" or k.startswith("utils.")]:
        sys.modules.pop(m)
    sys.modules.update(stash)
    return ref


def reference_affinity(ref, chunk, config_name):
    """Drive the reference's ncuts_chunk up to its normalized_cut call and return what it passes."""
    nu, cfg = ref["nu"], ref["cfg"]
    nu.CONFIG = getattr(cfg, "config_" + config_name)
    got = {}

    def capture(A, num_points, labels, T=None, split_lim=None):
        got.update(A=A, n=num_points, labels=labels, T=T, split_lim=split_lim)
        raise _Captured()

    def fake_tarl(dataset, chunk_major, T_pcd, center_position, tarl_indices_global):
        return chunk.tarl

    def fake_image(dataset, pcd_nonground_minor, chunk_indices, chunk_major, T_pcd, cam_indices_global,
                   sam=False, dino=False, pcd_chunk=None):
        assert dino and not sam
        return [chunk.dino[:, None, :]], None           # one camera, one view

    nu.normalized_cut = capture
    nu.tarl_features_per_patch = fake_tarl
    nu.image_based_features_per_patch = fake_image
    major = _FakeCloud(chunk.points)
    d = {
        "center_ids": [5], "center_positions": [np.zeros(3)], "indices": [np.arange(chunk.n)],
        "pcd_nonground_chunks": [major], "pcd_ground_chunks": [major],
        "pcd_nonground_chunks_major_downsampling": [major],
        "kitti_labels": {"ground": {"instance": [np.zeros(1)], "semantic": [np.zeros(1)]}},
    }
    try:
        nu.ncuts_chunk(None, d, None, np.eye(4), list(range(0, 40)), sequence=0, patchwise_indices=[[5]])
    except _Captured:
        pass
    else:
        raise RuntimeError("reference ncuts_chunk returned without calling normalized_cut")
    return got


def known_answer_cases():
    """Hand-checkable inputs: returns name -> dense float64 w."""
    cases = {}
    # two 6-cliques joined by one weak edge: the cut must separate them
    w = np.zeros((12, 12))
    w[:6, :6] = 0.9
    w[6:, 6:] = 0.8
    w[5, 6] = w[6, 5] = 0.01
    np.fill_diagonal(w, 1.0)
    cases["two_cliques"] = w
    # single clique: Fiedler space is degenerate, every cut is expensive -> one segment at T = 0.03
    w = np.full((8, 8), 0.7)
    np.fill_diagonal(w, 1.0)
    cases["one_clique"] = w
    # path graph with one weak link in the middle; unequal weights so that no Fiedler vector is
    # antisymmetric (a symmetric path makes the chosen cut depend on the arbitrary eigenvector sign)
    w = np.eye(10)
    for i, v in enumerate([0.6, 0.5, 0.7, 0.4, 0.05, 0.65, 0.45, 0.55, 0.75]):
        w[i, i + 1] = w[i + 1, i] = v
    cases["weak_path"] = w
    return cases


def main():
    sys.path.insert(0, REPO)
    from autoinst_b200.synthetic import CONFIGS, small_chunk
    from oracle import ncut_ref as R
    from oracle.affinity_ref import affinity_ref

    os.makedirs(OUT, exist_ok=True)
    ref = import_reference()
    ref_nc = ref["nc"].normalized_cut

    for seed, n_obj, ppo in [(11, 3, 250), (12, 4, 300), (13, 5, 350)]:
        ch = small_chunk(seed, n_obj=n_obj, pts_per_obj=ppo)
        np.savez_compressed(os.path.join(OUT, f"chunk_s{seed}_inputs.npz"), points=ch.points,
                            tarl=ch.tarl.astype(np.float32), dino=ch.dino.astype(np.float32), instance=ch.instance)
        for name, cfg in CONFIGS.items():
            got = reference_affinity(ref, ch, name)
            A_ref = got["A"].toarray()
            A_or = affinity_ref(ch.points, ch.tarl, ch.dino, alpha=cfg["alpha"], theta=cfg["theta"], gamma=cfg["gamma"])
            if not np.array_equal(A_ref, A_or):
                raise SystemExit(f"oracle affinity differs from the reference (seed {seed}, {name}): "
                                 f"max abs diff {np.abs(A_ref - A_or).max()}")
            assert got["n"] == ch.n and got["T"] == cfg["T"] and got["split_lim"] == 0.01
            w = got["A"]
            with R.pinned_eigsh():
                g_ref = ref_nc(w, ch.n, np.arange(ch.n), T=cfg["T"], split_lim=0.01)
                g_or = R.normalized_cut_ref(w, ch.n, np.arange(ch.n), T=cfg["T"], split_lim=0.01, faithful=True)
                g_fast = R.normalized_cut_ref(w, ch.n, np.arange(ch.n), T=cfg["T"], split_lim=0.01, faithful=False)
            with R.pinned_eigsh("random", 3):
                g_alt = ref_nc(w, ch.n, np.arange(ch.n), T=cfg["T"], split_lim=0.01)
            for g in (g_or, g_fast):
                if len(g) != len(g_ref) or any(not np.array_equal(a, b) for a, b in zip(g, g_ref)):
                    raise SystemExit(f"oracle normalized_cut differs from the reference (seed {seed}, {name})")
            lab = R.labels_from_groups(g_ref, ch.n)
            stable = R.same_partition(lab, R.labels_from_groups(g_alt, ch.n))
            csr = sp.csr_matrix(w)
            path = os.path.join(OUT, f"chunk_s{seed}_{name}.npz")
            np.savez_compressed(
                path, alpha=cfg["alpha"], theta=cfg["theta"], gamma=cfg["gamma"], T=cfg["T"],
                A_data=csr.data, A_indices=csr.indices, A_indptr=csr.indptr, labels=lab,
                group_sizes=np.array([len(g) for g in g_ref]), oracle_stable=stable)
            print(f"{os.path.basename(path)}: N={ch.n} nnz={csr.nnz} segments={len(g_ref)} stable={stable}")

    kat = {}
    for name, w in known_answer_cases().items():
        for T in (0.03, 0.5):
            with R.pinned_eigsh():
                g_ref = ref_nc(sp.csr_matrix(w), w.shape[0], np.arange(w.shape[0]), T=T, split_lim=0.01)
                g_or = R.normalized_cut_ref(sp.csr_matrix(w), w.shape[0], np.arange(w.shape[0]), T=T, split_lim=0.01)
            assert len(g_ref) == len(g_or) and all(np.array_equal(a, b) for a, b in zip(g_ref, g_or)), name
            with R.pinned_eigsh(flip=True):
                g_flip = ref_nc(sp.csr_matrix(w), w.shape[0], np.arange(w.shape[0]), T=T, split_lim=0.01)
            assert R.same_partition(R.labels_from_groups(g_ref, w.shape[0]), R.labels_from_groups(g_flip, w.shape[0])), \
                f"KAT {name} T={T} depends on the eigenvector sign"
            kat[f"{name}_T{T}_w"] = w
            kat[f"{name}_T{T}_labels"] = R.labels_from_groups(g_ref, w.shape[0])
            print(f"KAT {name} T={T}: {[list(map(int, g)) for g in g_ref]}")
    np.savez_compressed(os.path.join(OUT, "known_answers.npz"), **kat)
    golden_metrics()
    golden_pooling()
    golden_merge()
    golden_dino()


class _O3dCloud:
    """Stand-in for open3d.geometry.PointCloud as used by get_pcd / transform_pcd (point_cloud_utils.py:11-35)."""
    def __init__(self):
        self.points = np.zeros((0, 3))

    def transform(self, T):                       # Open3D: homogeneous transform of every point, in place, returns self
        P = np.asarray(self.points, dtype=np.float64)
        hom = np.concatenate([P, np.ones((P.shape[0], 1))], axis=1) @ np.asarray(T, dtype=np.float64).T
        self.points = hom[:, :3] / hom[:, 3:4]
        return self


class _O3dKDTree:
    """Stand-in for open3d.geometry.KDTreeFlann (nanoflann radiusSearch: dist^2 < radius^2, sorted by distance)."""
    def __init__(self, pcd):
        from scipy.spatial import cKDTree
        self.P = np.asarray(pcd.points, dtype=np.float64)
        self.tree = cKDTree(self.P) if self.P.shape[0] else None

    def search_radius_vector_3d(self, query, radius):
        if self.tree is None:
            return 0, np.zeros(0, dtype=np.int64), np.zeros(0)
        q = np.asarray(query, dtype=np.float64)
        cand = np.asarray(self.tree.query_ball_point(q, radius * (1.0 + 1e-9)), dtype=np.int64)
        d2 = ((self.P[cand] - q) ** 2).sum(axis=1) if cand.size else np.zeros(0)
        keep = d2 < radius * radius
        order = np.argsort(d2[keep], kind="stable")
        idx = cand[keep][order]
        return int(idx.size), idx, d2[keep][order]


def golden_pooling():
    """Run the reference's tarl_features_per_patch on synthetic scans and pin oracle.pool_ref against it."""
    from autoinst_b200.synthetic import small_chunk, make_scans
    from oracle.pool_ref import pool_features_ref
    for name in ["open3d", "open3d.geometry", "open3d.utility", "open3d.io", "open3d.pipelines",
                 "open3d.pipelines.registration"]:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    o3d = sys.modules["open3d"]
    o3d.geometry = sys.modules["open3d.geometry"]
    o3d.utility = sys.modules["open3d.utility"]
    o3d.pipelines = sys.modules["open3d.pipelines"]
    o3d.pipelines.registration = sys.modules["open3d.pipelines.registration"]
    o3d.geometry.PointCloud = _O3dCloud
    o3d.geometry.KDTreeFlann = _O3dKDTree
    o3d.utility.Vector3dVector = lambda a: np.asarray(a, dtype=np.float64)
    cwd = os.getcwd()
    os.chdir(REF)
    sys.path.insert(0, REF)
    stash = {m: sys.modules.pop(m) for m in [k for k in sys.modules if k == "config" or k == "utils" or k.startswith("utils.")]}
    try:
        import utils.point_cloud.chunk_generation as cg
        assert cg.__file__.startswith(REF)
    finally:
        os.chdir(cwd)
        sys.path.remove(REF)
    ch = small_chunk(21, n_obj=2, pts_per_obj=160, features="tarl")
    scans = make_scans(ch, n_scans=3, pts_per_major=1.5, seed=3)

    def pose(k):
        a = 0.2 * k
        T = np.eye(4)
        T[:3, :3] = [[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]]
        T[:3, 3] = [1.5 * k, -0.5 * k, 0.05 * k]
        return T
    T_pcd = pose(2)
    lidar = []                                   # every scan in its own lidar frame; the reference brings it back (:228-231)
    for k, (pts, _) in enumerate(scans):
        Tinv = np.linalg.inv(np.linalg.inv(T_pcd) @ pose(k))
        lidar.append(pts @ Tinv[:3, :3].T + Tinv[:3, 3])

    class Dataset:
        def get_pose(self, i): return pose(i)
        def get_point_cloud(self, i): return lidar[i]
        def get_tarl_features(self, i): return scans[i][1]
    pcd = _O3dCloud()
    pcd.points = ch.points
    ref_out = cg.tarl_features_per_patch(Dataset(), pcd, T_pcd, ch.center, list(range(len(scans))))
    # the oracle gets the scans as the reference sees them after its own transform (:231)
    seen = []
    for k in range(len(scans)):
        c = _O3dCloud()
        c.points = lidar[k]
        seen.append((c.transform(np.linalg.inv(T_pcd) @ pose(k)).points, scans[k][1]))
    mine, cnt = pool_features_ref(ch.points, seen, ch.center, radius=cg.MAJOR_VOXEL_SIZE / 2., chunk_size=cg.CHUNK_SIZE,
                                  normalise=cg.TARL_NORM, return_count=True)
    if not np.array_equal(ref_out, mine):
        raise SystemExit(f"oracle pooling differs from the reference: max abs diff {np.abs(ref_out - mine).max()}")
    assert (cnt == 0).any() and (cnt > 3).any()
    np.savez_compressed(os.path.join(OUT, "pooling.npz"), major=ch.points, center=ch.center,
                        scan_points=np.concatenate([s[0] for s in seen]), scan_features=np.concatenate([s[1] for s in seen]),
                        scan_sizes=np.array([s[0].shape[0] for s in seen]), radius=cg.MAJOR_VOXEL_SIZE / 2.,
                        chunk_size=np.asarray(cg.CHUNK_SIZE, dtype=np.float64), out=ref_out, count=cnt)
    print(f"pooling.npz: {ch.n} major points, {sum(s[0].shape[0] for s in seen)} scan points, zero rows {(cnt == 0).sum()}")
    for m_ in [k for k in sys.modules if k == "config" or k == "utils" or k.startswith("utils.")]:
        sys.modules.pop(m_)
    sys.modules.update(stash)


class _O3dColourCloud(_O3dCloud):
    """PointCloud stand-in with colours for the merge (`point_cloud_utils.py:320-329,387-491`): `+=`, `crop` with an
    inclusive axis-aligned box (Open3D `GetPointIndicesWithinBoundingBox`), `remove_duplicated_points` keeping the first
    occurrence of every exact coordinate in point order."""
    def __init__(self, points=None, colors=None):
        self.points = np.zeros((0, 3)) if points is None else np.asarray(points, dtype=np.float64)
        self.colors = np.zeros((len(self.points), 3)) if colors is None else np.asarray(colors, dtype=np.float64)

    def __iadd__(self, other):
        self.points = np.concatenate([np.asarray(self.points), np.asarray(other.points)])
        self.colors = np.concatenate([np.asarray(self.colors), np.asarray(other.colors)])
        return self

    def crop(self, box):
        P = np.asarray(self.points)
        m = np.all((P >= box.min_bound) & (P <= box.max_bound), axis=1)
        return _O3dColourCloud(P[m], np.asarray(self.colors)[m])

    def remove_duplicated_points(self):
        _, first = np.unique(np.asarray(self.points), axis=0, return_index=True)
        keep = np.sort(first)
        self.points = np.asarray(self.points)[keep]
        self.colors = np.asarray(self.colors)[keep]
        return self


class _O3dBox:
    def __init__(self, min_bound=None, max_bound=None):
        self.min_bound = np.asarray(min_bound, dtype=np.float64)
        self.max_bound = np.asarray(max_bound, dtype=np.float64)


def golden_merge():
    """Run the reference's merge_chunks_unite_instances2 / merge_unite_gt / remove_semantics on a synthetic map and pin
    oracle.merge_ref against them (labels <-> colours: label l is the colour (l / 4096, 0, 0), black = background, so the
    reference's np.unique order of colours is the order of the labels)."""
    from autoinst_b200.synthetic import make_map
    from oracle import merge_ref as M
    for name in ["open3d", "open3d.geometry", "open3d.utility", "open3d.io", "open3d.pipelines",
                 "open3d.pipelines.registration"]:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    o3d = sys.modules["open3d"]
    o3d.geometry = sys.modules["open3d.geometry"]
    o3d.utility = sys.modules["open3d.utility"]
    o3d.pipelines = sys.modules["open3d.pipelines"]
    o3d.pipelines.registration = sys.modules["open3d.pipelines.registration"]
    o3d.geometry.PointCloud = _O3dColourCloud
    o3d.geometry.AxisAlignedBoundingBox = _O3dBox
    o3d.utility.Vector3dVector = lambda a: np.asarray(a, dtype=np.float64)
    cwd = os.getcwd()
    os.chdir(REF)
    sys.path.insert(0, REF)
    stash = {m: sys.modules.pop(m) for m in [k for k in sys.modules if k == "config" or k == "utils" or k.startswith("utils.")]}
    try:
        import utils.point_cloud.point_cloud_utils as pcu
        assert pcu.__file__.startswith(REF)
    finally:
        os.chdir(cwd)
        sys.path.remove(REF)
    chunks = [c for c in make_map(4, 1500, seed=5) if c.n > 0]
    rng = np.random.default_rng(1)
    parts = []
    for c in chunks:
        # imperfect "predictions": GT instances with one of them split in two and some points on background
        lab = c.instance.astype(np.int64).copy()
        if (lab > 0).any():
            big = np.bincount(lab[lab > 0]).argmax()
            half = (lab == big) & (c.points[:, 0] > np.median(c.points[lab == big, 0]))
            lab[half] = lab.max() + 1
        lab[rng.random(c.n) < 0.03] = 0
        lab = np.where(lab > 0, (np.int64(c.chunk_id + 1) << 6) + lab, 0)            # unique across chunks, < 4096
        parts.append((c.points, lab))
    assert max(int(l.max()) for _, l in parts) < 4096

    def cloud(p, l):
        col = np.zeros((len(l), 3))
        col[:, 0] = l / 4096.0
        return _O3dColourCloud(p, col)
    merged = pcu.merge_chunks_unite_instances2([cloud(p, l) for p, l in parts])
    ref_pts = np.asarray(merged.points)
    ref_lab = np.rint(np.asarray(merged.colors)[:, 0] * 4096.0).astype(np.int64)
    pts, lab = M.merge_chunks_unite_instances(parts)
    if not (np.array_equal(pts, ref_pts) and np.array_equal(lab, ref_lab)):
        raise SystemExit("oracle merge differs from the reference's merge_chunks_unite_instances2")
    gt = pcu.merge_unite_gt([cloud(c.points, c.instance.astype(np.int64)) for c in chunks])
    gpts, glab = M.merge_unite_gt([(c.points, c.instance) for c in chunks])
    if not (np.array_equal(gpts, np.asarray(gt.points)) and
            np.array_equal(glab, np.rint(np.asarray(gt.colors)[:, 0] * 4096.0).astype(np.int64))):
        raise SystemExit("oracle merge_unite_gt differs from the reference")
    assert np.array_equal(gpts, pts)
    pred = M.compact_labels(lab)
    ref_clean = pcu.remove_semantics(M.compact_labels(glab), pred.copy())
    mine_clean = M.remove_semantics(M.compact_labels(glab), pred)
    if not np.array_equal(ref_clean, mine_clean):
        raise SystemExit("oracle remove_semantics differs from the reference")
    np.savez_compressed(os.path.join(OUT, "merge.npz"), n_chunks=len(parts),
                        **{f"p{i}": p for i, (p, _) in enumerate(parts)}, **{f"l{i}": l for i, (_, l) in enumerate(parts)},
                        **{f"g{i}": c.instance.astype(np.int64) for i, c in enumerate(chunks)},
                        merged_points=ref_pts, merged_labels=ref_lab, gt_labels=glab, cleaned=ref_clean)
    print(f"merge.npz: {len(parts)} chunks, {sum(len(l) for _, l in parts)} points -> {len(ref_lab)} merged, "
          f"{len(np.unique(ref_lab)) - 1} instances (from {sum(len(np.unique(l)) - 1 for _, l in parts)} per-chunk segments)")
    for m_ in [k for k in sys.modules if k == "config" or k == "utils" or k.startswith("utils.")]:
        sys.modules.pop(m_)
    sys.modules.update(stash)


def golden_metrics():
    """Run the reference's own Metrics class (`pipeline/metrics/metrics_class.py`) on seeded label arrays and
    check oracle.metrics_ref against it; store inputs + reference outputs in tests/golden/metrics.npz."""
    from oracle.metrics_ref import instance_metrics
    cwd = os.getcwd()
    os.chdir(REF)
    sys.path.insert(0, REF)
    stash = {m: sys.modules.pop(m) for m in [k for k in sys.modules if k == "config" or k == "metrics" or k.startswith("metrics.")]}
    try:
        import metrics.metrics_class as mc
    finally:
        os.chdir(cwd)
        sys.path.remove(REF)
    out = {}
    rng = np.random.default_rng(42)
    for case in range(4):
        n = 6000
        n_gt = 6 + case
        gt = rng.integers(0, n_gt + 1, size=n)
        gt = np.sort(gt)                                        # contiguous instances
        pred = gt.copy()
        # perturb: split some instances, merge others, add noise and a background-heavy prediction
        pred[(gt == 2) & (rng.random(n) < 0.5)] = n_gt + 5
        pred[gt == 4] = 3
        noise = rng.random(n) < (0.03 + 0.02 * case)
        pred[noise] = rng.integers(0, n_gt + 8, size=int(noise.sum()))
        allp = pred.copy()
        pred2 = pred.copy()
        pred2[pred2 == 1] = 0                                   # what remove_semantics would do
        for mp in (200, 20):
            m = mc.Metrics(name=f"golden{case}", min_points=mp)
            ref_out, ref_ap = m.update_stats(allp.copy(), pred2.copy(), gt.copy())
            ref = {"p": ref_out["precision"], "r": ref_out["recall"], "f1": ref_out["fScore"], "ap": ref_ap["ap"],
                   "ap0.25": ref_ap["0.25"], "ap0.5": ref_ap["0.5"], "S_assoc": ref_ap["lstq"]}
            mine = instance_metrics(allp, pred2, gt, min_points=mp)
            for k in ref:
                if not np.isclose(ref[k], mine[k], rtol=0, atol=1e-12):
                    raise SystemExit(f"oracle metrics differ from the reference: case {case} min_points {mp} {k}: {ref[k]} vs {mine[k]}")
            out[f"c{case}_mp{mp}_ref"] = np.array([ref[k] for k in ("p", "r", "f1", "ap", "ap0.25", "ap0.5", "S_assoc")])
            print(f"metrics case {case} min_points {mp}:", {k: round(float(v), 4) for k, v in ref.items()})
        out[f"c{case}_all"] = allp
        out[f"c{case}_pred"] = pred2
        out[f"c{case}_gt"] = gt
    for m_ in [k for k in sys.modules if k == "config" or k == "metrics" or k.startswith("metrics.")]:
        sys.modules.pop(m_)
    sys.modules.update(stash)
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), **out)


if __name__ == "__main__":
    main()


# ---------------------------------------------------------------------------------------------------------------------
# row N3, DINOv2 half: image_based_features_per_patch (sam=False, dino=True, hpr_masks given) + dinov2_mean
# ---------------------------------------------------------------------------------------------------------------------
class _O3dCloudFull(_O3dCloud):
    """PointCloud stand-in for image_utils.py: points, transform, statistical outlier removal (stand-in: k-NN mean
    distance against mean + std_ratio * std, Open3D's published rule) — the outlier filter is NOT part of the restated
    path, it only has to give the reference function a deterministic index list."""
    def remove_statistical_outlier(self, nb_neighbors=20, std_ratio=2.0):
        from scipy.spatial import cKDTree
        P = np.asarray(self.points, dtype=np.float64)
        k = min(nb_neighbors, P.shape[0])
        d, _ = cKDTree(P).query(P, k=k)
        avg = d.reshape(P.shape[0], -1).mean(axis=1)
        keep = np.where(avg < avg.mean() + std_ratio * avg.std())[0]
        return None, keep


class _O3dKDTreeKnn(_O3dKDTree):
    def search_knn_vector_3d(self, query, k):
        d, idx = self.tree.query(np.asarray(query, dtype=np.float64), k=k)
        idx = np.atleast_1d(idx)
        d = np.atleast_1d(d)
        return int(idx.size), idx, d * d


def golden_dino():
    """Run the reference's image_based_features_per_patch + dinov2_mean on a synthetic scene; oracle.dino_ref must match."""
    from autoinst_b200.synthetic import small_chunk
    from oracle.dino_ref import dino_mean_ref, transform_points
    for name in ["open3d", "open3d.geometry", "open3d.utility", "open3d.io", "open3d.pipelines",
                 "open3d.pipelines.registration", "matplotlib", "matplotlib.pyplot"]:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    o3d = sys.modules["open3d"]
    o3d.geometry = sys.modules["open3d.geometry"]
    o3d.utility = sys.modules["open3d.utility"]
    o3d.pipelines = sys.modules["open3d.pipelines"]
    o3d.pipelines.registration = sys.modules["open3d.pipelines.registration"]
    o3d.geometry.PointCloud = _O3dCloudFull
    o3d.geometry.KDTreeFlann = _O3dKDTreeKnn
    o3d.geometry.KDTreeSearchParamHybrid = lambda **kw: None
    o3d.utility.Vector3dVector = lambda a: np.asarray(a, dtype=np.float64)
    plt = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].pyplot = plt
    if not hasattr(plt, "cm"):
        plt.cm = types.SimpleNamespace(viridis=lambda x: np.zeros((len(x), 4)))
    cwd = os.getcwd()
    os.chdir(REF)
    sys.path.insert(0, REF)
    stash = {m: sys.modules.pop(m) for m in [k for k in sys.modules if k == "config" or k == "utils" or k.startswith("utils.")]}
    try:
        import utils.image.image_utils as iu
        assert iu.__file__.startswith(REF)
    finally:
        os.chdir(cwd)
        sys.path.remove(REF)
    from autoinst_b200.synthetic import make_camera_scene
    sc = make_camera_scene(31)
    major, pcd_pts, chunk_indices, T_pcd2world = sc["major"], sc["pcd_points"], sc["chunk_indices"], sc["T_pcd2world"]
    n, n_views, hpr_masks, K, T_lidar2cam, fmaps, pose = major.shape[0], len(sc["cam_indices"]), sc["hpr_masks"], sc["K"], \
        sc["T_lidar2cam"], sc["feature_maps"], sc["pose"]
    img_h, img_w = sc["img_hw"]
    Dataset = lambda: sc["dataset"]      # noqa: E731
    pcd = _O3dCloudFull()
    pcd.points = pcd_pts
    chunk_nc = _O3dCloudFull()
    chunk_nc.points = major
    cam_indices = list(range(n_views))
    lists, _vis = iu.image_based_features_per_patch(Dataset(), pcd, chunk_indices, chunk_nc, T_pcd2world, cam_indices,
                                                    hpr_masks=hpr_masks, sam=False, dino=True)
    assert len(lists) == 1
    ref = iu.dinov2_mean(lists[0])
    # the same views as plain arrays: what the restated path (and the CUDA path) takes
    sub = _O3dCloudFull()
    sub.points = pcd_pts[chunk_indices]
    _, inl = sub.remove_statistical_outlier(nb_neighbors=20, std_ratio=2.0)
    chunk_and_inlier = set(chunk_indices[inl].tolist())
    views = []
    for i in cam_indices:
        T_pcd2cam = T_lidar2cam @ np.linalg.inv(pose(i)) @ T_pcd2world                  # image_utils.py:150-154
        frame = list(set(np.where(hpr_masks[i])[0].tolist()) & chunk_and_inlier)        # :206-209 (the reference's own order)
        if len(frame) == 0:
            views.append(None)
            continue
        views.append(dict(T_pcd2cam=T_pcd2cam, visible_cam=transform_points(pcd_pts, T_pcd2cam)[frame], K=K,
                          img_hw=(img_h, img_w), feature_map=fmaps[i]))
    mine = dino_mean_ref(major, views, max_dist=iu.MAJOR_VOXEL_SIZE / 2, fdim=iu.NUM_DINO_FEATURES)
    if not np.array_equal(ref, mine):
        raise SystemExit(f"oracle DINO mean differs from the reference: max abs diff {np.abs(ref - mine).max()}")
    seen = np.count_nonzero(ref.any(axis=1))
    assert 0 < seen < n and views[3] is None
    out = dict(major=major, n_views=np.int64(n_views), K=K, img_hw=np.array([img_h, img_w]), out=ref,
               max_dist=np.float64(iu.MAJOR_VOXEL_SIZE / 2))
    for i, v in enumerate(views):
        out[f"skip{i}"] = np.bool_(v is None)
        if v is not None:
            out[f"T{i}"] = v["T_pcd2cam"]; out[f"vis{i}"] = v["visible_cam"]; out[f"fmap{i}"] = v["feature_map"]
    np.savez_compressed(os.path.join(OUT, "dino.npz"), **out)
    print(f"dino.npz: {n} major points, {n_views} views, {seen} points with features")
    for m_ in [k for k in sys.modules if k == "config" or k == "utils" or k.startswith("utils.")]:
        sys.modules.pop(m_)
    sys.modules.update(stash)
