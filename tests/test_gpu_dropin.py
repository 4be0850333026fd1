"""GPU tests of the reference-facing call surface: `ncuts.normalized_cut.normalized_cut` and
`ncuts.ncuts_utils.ncuts_chunk` with the reference's signatures (run_pipeline.py:14-17,165-180)."""
import sys
import types

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import CONFIG_NAMES, GOLDEN_SEEDS, load_golden
from autoinst_b200.synthetic import CONFIGS, make_chunk, make_scans
from oracle import ncut_ref as R
from oracle.affinity_ref import affinity_ref
from oracle.pool_ref import pool_features_ref

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.mark.parametrize("seed", GOLDEN_SEEDS[:2])
@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_normalized_cut_signature_and_result(cuda_device, seed, name):
    from ncuts.normalized_cut import normalized_cut
    inp, out, A = load_golden(seed, name)
    n = A.shape[0]
    groups = normalized_cut(A, n, np.arange(n), T=float(out["T"]), split_lim=0.01)       # scipy CSR in, list of arrays out
    assert isinstance(groups, list) and sorted(np.concatenate(groups).tolist()) == list(range(n))
    assert R.same_partition(R.labels_from_groups(groups, n), out["labels"])
    # a sub-block with its own labels and the chunk's num_points_orig, as the reference's recursion calls it (:57-58)
    idx = np.where(out["labels"] == 0)[0]
    sub = normalized_cut(A[idx][:, idx], n, idx, T=float(out["T"]))
    with R.pinned_eigsh():
        ref = R.normalized_cut_ref(A[idx][:, idx], n, idx, T=float(out["T"]))
    assert sorted(map(lambda g: tuple(sorted(g.tolist())), sub)) == sorted(map(lambda g: tuple(sorted(g.tolist())), ref))


class FakeCloud:
    """Just enough of open3d.geometry.PointCloud for ncuts_chunk."""

    def __init__(self, points, colors=None):
        self.points = np.asarray(points, dtype=np.float64)
        self.colors = np.zeros_like(self.points) if colors is None else np.asarray(colors, dtype=np.float64)

    def paint_uniform_color(self, c):
        self.colors = np.tile(np.asarray(c, dtype=np.float64), (len(self.points), 1))

    def select_by_index(self, idx):
        idx = np.asarray(idx)
        return FakeCloud(self.points[idx], self.colors[idx])

    def __add__(self, other):
        return FakeCloud(np.concatenate([self.points, other.points]), np.concatenate([self.colors, other.colors]))


def install_fake_reference_modules(monkeypatch, chunk):
    """Stub the reference modules ncuts_chunk imports lazily (open3d, utils.*): feature fetchers return the
    synthetic arrays, geometry helpers are numpy one-liners."""
    o3d = types.ModuleType("open3d")
    o3d.utility = types.SimpleNamespace(Vector3dVector=lambda a: np.asarray(a))
    o3d.geometry = types.SimpleNamespace(PointCloud=FakeCloud)
    pcu = types.ModuleType("utils.point_cloud.point_cloud_utils")
    pcu.get_subpcd = lambda pcd, idx: pcd.select_by_index(idx)
    pcu.get_statistical_inlier_indices = lambda pcd, nb_neighbors=20, std_ratio=2.0: np.arange(len(pcd.points))
    vis = types.ModuleType("utils.visualization_utils")
    vis.generate_random_colors = lambda n: [((37 * i) % 255 + 1, (91 * i) % 256, (53 * i) % 256) for i in range(n)]
    # the DINOv2 means come from autoinst_b200.dino (tests/test_dino.py checks that path against the reference's own
    # functions); here the per-camera result is the synthetic array
    import autoinst_b200.dino as dino_mod
    monkeypatch.setattr(dino_mod, "dinov2_mean_per_patch", lambda *a, **k: [chunk.dino])
    cg = types.ModuleType("utils.point_cloud.chunk_generation")      # tarl_features_per_patch is autoinst_b200.pooling's now

    def get_indices_feature_reprojection(global_indices, first_id, adjacent_frames=(8, 5)):
        i = global_indices.index(first_id)
        sel = global_indices[max(0, i - adjacent_frames[0]): i + adjacent_frames[1]]
        return sel, [global_indices.index(g) for g in sel]
    cg.get_indices_feature_reprojection = get_indices_feature_reprojection
    for name, mod in {"open3d": o3d, "utils": types.ModuleType("utils"), "utils.point_cloud": types.ModuleType("utils.point_cloud"),
                      "utils.image": types.ModuleType("utils.image"), "utils.point_cloud.point_cloud_utils": pcu,
                      "utils.visualization_utils": vis,
                      "utils.point_cloud.chunk_generation": cg}.items():
        monkeypatch.setitem(sys.modules, name, mod)


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_ncuts_chunk_drop_in(cuda_device, monkeypatch, name):
    import ncuts.ncuts_utils as nu
    cfg = dict(CONFIGS[name], name=name, out_folder="x/", gt=True)
    ch = make_chunk(31, n_target=1500, features="tarl_dino")
    install_fake_reference_modules(monkeypatch, ch)
    monkeypatch.setattr(nu, "CONFIG", cfg)                               # the way the survey switches configs (§5)
    rng = np.random.default_rng(0)
    minor = ch.points[rng.integers(0, ch.n, size=6000)] + rng.normal(0, 0.03, size=(6000, 3))     # 5 cm cloud
    ground = np.stack([rng.uniform(-12, 12, 500), rng.uniform(-12, 12, 500), rng.normal(-12.4, 0.05, 500)], 1)
    d = {"center_ids": [5], "center_positions": [np.zeros(3)], "indices": [np.arange(len(minor))],
         "pcd_nonground_chunks": [FakeCloud(minor)], "pcd_ground_chunks": [FakeCloud(ground)],
         "pcd_nonground_chunks_major_downsampling": [FakeCloud(ch.points)],
         "kitti_labels": {"ground": {"instance": [np.zeros(500, int)], "semantic": [np.full(500, 40)]}}}
    # TARL inputs come per scan from the dataset and are pooled onto the major points on the GPU
    # (chunk_generation.py:205-258 -> autoinst_b200.pooling); scans 4, 5, 6 carry points, the others are empty
    scans = dict(zip((4, 5, 6), make_scans(ch, n_scans=3, pts_per_major=2.0)))
    none = (np.zeros((0, 3)), np.zeros((0, 96), dtype=np.float32))

    class FakeDataset:
        def get_pose(self, i): return np.eye(4)
        def get_point_cloud(self, i): return scans.get(i, none)[0]
        def get_tarl_features(self, i): return scans.get(i, none)[1]
    merged, pcd_chunk, cut_hight, inst_g, seg_g = nu.ncuts_chunk(FakeDataset(), d, None, np.eye(4), list(range(40)),
                                                                 sequence=0, patchwise_indices=[[5]])
    # the five return values of the reference (ncuts_utils.py:204)
    assert len(merged.points) == len(minor) + len(cut_hight.points) and len(inst_g) == len(seg_g) == len(cut_hight.points)
    assert np.all(np.asarray(cut_hight.colors) == 0)                     # ground painted black
    # colours encode the segments: decode and compare with the oracle labels re-projected by nearest neighbour
    tarl_ref = pool_features_ref(ch.points, [scans[i] for i in (4, 5, 6)], np.zeros(3))
    assert (~tarl_ref.any(axis=1)).any()                                 # some rows have no scan point (no_tarl_mask)
    A = affinity_ref(ch.points, tarl_ref, ch.dino, alpha=cfg["alpha"], theta=cfg["theta"], gamma=cfg["gamma"])
    with R.pinned_eigsh():
        g = R.normalized_cut_ref(sp.csr_matrix(A), ch.n, np.arange(ch.n), T=cfg["T"])
    ref_major = R.labels_from_groups(g, ch.n)
    from scipy.spatial import cKDTree
    _, nn = cKDTree(ch.points).query(minor, k=1)
    _, got = np.unique(np.asarray(pcd_chunk.colors), axis=0, return_inverse=True)
    assert R.same_partition(got.reshape(-1), ref_major[nn])
    with pytest.raises(ValueError):
        monkeypatch.setattr(nu, "CONFIG", dict(cfg, gamma=0.1))
        import autoinst_b200.dino as dino_mod
        monkeypatch.setattr(dino_mod, "dinov2_mean_per_patch", lambda *a, **k: [])
        nu.ncuts_chunk(FakeDataset(), d, None, np.eye(4), list(range(40)), sequence=0, patchwise_indices=[[5]])
