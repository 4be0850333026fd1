"""Row N3, DINOv2 half (per-view feature look-up + mean over views, image_utils.py:264-346,363-371): the oracle against the
golden vector written by the reference's own functions (CPU), the CUDA path and its host mirror against both (GPU)."""
import sys
import types

import numpy as np
import pytest

from conftest import GOLDEN
from autoinst_b200.synthetic import make_camera_scene
from oracle.dino_ref import dino_mean_ref, transform_points, view_pixels_ref


def golden_views():
    g = np.load(f"{GOLDEN}/dino.npz")
    views = []
    for i in range(int(g["n_views"])):
        if bool(g[f"skip{i}"]):
            views.append(None)
        else:
            views.append(dict(T_pcd2cam=g[f"T{i}"], visible_cam=g[f"vis{i}"], K=g["K"], img_hw=tuple(int(x) for x in g["img_hw"]),
                              feature_map=g[f"fmap{i}"]))
    return g, views


def test_oracle_matches_the_reference_function_golden():
    """tests/golden/dino.npz: image_based_features_per_patch(sam=False, dino=True, hpr_masks=...) + dinov2_mean, run
    unmodified by oracle/make_golden.py::golden_dino."""
    g, views = golden_views()
    out = dino_mean_ref(g["major"], views, max_dist=float(g["max_dist"]))
    assert np.array_equal(out, g["out"])
    seen = out.any(axis=1)
    assert 0 < seen.sum() < len(seen) and views[3] is None


def test_oracle_hand_cases():
    K = np.array([[100.0, 0, 50.0], [0, 100.0, 40.0], [0, 0, 1.0]])
    major = np.array([[0.0, 0.0, 2.0], [0.5, 0.0, 2.0], [0.0, 0.0, -2.0], [5.0, 0.0, 2.0], [0.1, 0.1, 2.0]])
    vis = np.array([[0.0, 0.0, 2.1], [0.5, 0.0, 2.3], [0.0, 0.0, -2.0], [5.0, 0.0, 2.0], [0.1, 0.1, 2.0]])
    pix = view_pixels_ref(major, vis, K, 80, 100, 8, 10, max_dist=0.175)
    # point 0: 0.1 m from its neighbour -> pixel (50, 40) -> map (4, 5); point 1: 0.3 m away -> dropped; point 2: behind the
    # camera (z <= 0); point 3: outside the image (u = 300); point 4: pixel (55, 45) -> map (4, 5)
    assert pix.tolist() == [[4, 5], [-1, -1], [-1, -1], [-1, -1], [4, 5]]
    fmap = np.zeros((8, 10, 4), dtype=np.float32)
    fmap[4, 5] = [1, 2, 3, 4]
    zero = np.zeros((8, 10, 4), dtype=np.float32)
    v = dict(T_pcd2cam=np.eye(4), visible_cam=vis, K=K, img_hw=(80, 100), feature_map=fmap)
    out = dino_mean_ref(major, [v, None, dict(v, feature_map=zero), dict(v, feature_map=3 * fmap)], max_dist=0.175, fdim=4)
    assert np.array_equal(out[0], [2, 4, 6, 8]) and np.array_equal(out[4], [2, 4, 6, 8])      # views 0 and 3 count, the zero map not
    assert not out[1:4].any()


@pytest.mark.gpu
def test_dino_mean_matches_the_reference_golden(cuda_device):
    from autoinst_b200 import api
    g, views = golden_views()
    out, cnt = api.dino_mean_views(g["major"], views, float(g["max_dist"]), device=cuda_device, return_count=True)
    out = out.cpu().numpy()
    assert np.array_equal(out != 0, g["out"] != 0)
    assert np.array_equal(out, g["out"])                       # same pixels, float64 sums in view order: bit for bit
    assert cnt.cpu().numpy().max() <= 4 and (cnt.cpu().numpy() == 0).any()


@pytest.mark.gpu
def test_dino_edge_cases(cuda_device):
    from autoinst_b200 import api
    K = np.array([[100.0, 0, 50.0], [0, 100.0, 40.0], [0, 0, 1.0]])
    major = np.array([[0.0, 0.0, 2.0], [0.5, 0.0, 2.0], [0.0, 0.0, -2.0], [5.0, 0.0, 2.0], [0.1, 0.1, 2.0]])
    vis = np.array([[0.0, 0.0, 2.1], [0.5, 0.0, 2.3], [0.0, 0.0, -2.0], [5.0, 0.0, 2.0], [0.1, 0.1, 2.0]])
    fmap = np.zeros((8, 10, 4), dtype=np.float32)
    fmap[4, 5] = [1, 2, 3, 4]
    v = dict(T_pcd2cam=np.eye(4), visible_cam=vis, K=K, img_hw=(80, 100), feature_map=fmap)
    views = [v, None, dict(v, feature_map=np.zeros_like(fmap)), dict(v, feature_map=3 * fmap), dict(v, visible_cam=np.zeros((0, 3)))]
    out = api.dino_mean_views(major, views, 0.175, feat_dim=4, device=cuda_device).cpu().numpy()
    assert np.array_equal(out, dino_mean_ref(major, views, max_dist=0.175, fdim=4))
    # no view at all: zero rows
    assert not api.dino_mean_views(major, [], 0.175, feat_dim=4, device=cuda_device).cpu().numpy().any()
    assert not api.dino_mean_views(major, [None, None], 0.175, feat_dim=4, device=cuda_device).cpu().numpy().any()


@pytest.mark.gpu
def test_dinov2_mean_per_patch_drop_in(cuda_device, monkeypatch):
    """The host mirror with the reference's argument list (dataset, pcd, chunk_indices, chunk_nc, T_pcd2world, cam_indices,
    hpr_masks) on the scene the golden vector was made from: it must give the golden means."""
    sc = make_camera_scene(31)

    class Cloud:
        def __init__(self, pts): self.points = np.asarray(pts, dtype=np.float64)

    pcu = types.ModuleType("utils.point_cloud.point_cloud_utils")
    pcu.get_subpcd = lambda pcd, idx: Cloud(np.asarray(pcd.points)[np.asarray(idx)])

    def inliers(pcd, nb_neighbors=20, std_ratio=2.0):             # the stand-in of oracle/make_golden.py::_O3dCloudFull
        from scipy.spatial import cKDTree
        P = np.asarray(pcd.points)
        d, _ = cKDTree(P).query(P, k=min(nb_neighbors, P.shape[0]))
        avg = d.reshape(P.shape[0], -1).mean(axis=1)
        return np.where(avg < avg.mean() + std_ratio * avg.std())[0]
    pcu.get_statistical_inlier_indices = inliers
    for name, mod in {"utils": types.ModuleType("utils"), "utils.point_cloud": types.ModuleType("utils.point_cloud"),
                      "utils.point_cloud.point_cloud_utils": pcu}.items():
        monkeypatch.setitem(sys.modules, name, mod)
    from autoinst_b200.dino import dinov2_mean_per_patch
    out = dinov2_mean_per_patch(sc["dataset"], Cloud(sc["pcd_points"]), sc["chunk_indices"], Cloud(sc["major"]), sc["T_pcd2world"],
                                sc["cam_indices"], hpr_masks=sc["hpr_masks"], device=cuda_device)
    g = np.load(f"{GOLDEN}/dino.npz")
    assert len(out) == 1 and np.array_equal(out[0], g["out"])
