"""Feature pooling row (`chunk_generation.py:205-258`): the oracle on hand-checkable cases (CPU)
and the CUDA path (`ancuts_feature_pool`, through the C ABI) against the oracle (GPU)."""
import ctypes

import numpy as np
import pytest

from autoinst_b200.synthetic import make_chunk, make_scans
from oracle.pool_ref import crop_scan, pool_features_ref

R = 0.35 / 2.0


def hand_case():
    major = np.array([[0.0, 0.0, 0.0], [5.0, 0.0, 0.0], [12.0, 12.0, 12.0]])
    pts = np.array([[0.1, 0.0, 0.0],            # inside the radius of major 0
                    [0.0, R, 0.0],              # exactly on the sphere: excluded (strict <)
                    [0.0, 0.0, -0.17],          # inside
                    [5.0, 0.1, 0.1],            # inside the radius of major 1
                    [12.0, 12.0, 12.4],         # out of range of major 2
                    [12.5, 12.0, 12.0],         # on the face of the cube: cropped (strict >, <)
                    [13.0, 12.0, 12.05]])       # outside the cube, although within R of nothing anyway
    feats = np.arange(7 * 4, dtype=np.float32).reshape(7, 4)
    return major, pts, feats


def test_oracle_hand_case():
    major, pts, feats = hand_case()
    out, cnt = pool_features_ref(major, [(pts, feats)], np.zeros(3), radius=R, return_count=True)
    assert cnt.tolist() == [2, 1, 0]
    assert np.array_equal(out[0], (feats[0].astype(np.float64) + feats[2]) / 2)
    assert np.array_equal(out[1], feats[3].astype(np.float64))
    assert not out[2].any()                                    # zero row: neutralised later, ncuts_utils.py:143-146
    c, f = crop_scan(pts, feats, np.zeros(3))
    assert c.shape[0] == 5                                     # the face point and the outside point are dropped
    outn = pool_features_ref(major, [(pts, feats)], np.zeros(3), radius=R, normalise=True)
    assert abs(np.linalg.norm(outn[0]) - 1.0) < 1e-15 and not outn[2].any()


def load_pooling_golden():
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "pooling.npz"))
    sizes = g["scan_sizes"]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    scans = [(g["scan_points"][a:b], g["scan_features"][a:b]) for a, b in zip(offs[:-1], offs[1:])]
    return g, scans


def test_oracle_matches_the_reference_function_golden():
    """tests/golden/pooling.npz holds the output of the reference's own `tarl_features_per_patch`
    (`chunk_generation.py:205-258`, run unmodified by oracle/make_golden.py with a cKDTree stand-in for Open3D's KD-tree)."""
    g, scans = load_pooling_golden()
    out, cnt = pool_features_ref(g["major"], scans, g["center"], radius=float(g["radius"]), chunk_size=g["chunk_size"],
                                 return_count=True)
    assert np.array_equal(out, g["out"]) and np.array_equal(cnt, g["count"])
    assert (cnt == 0).any() and not out[cnt == 0].any()


def test_oracle_scans_are_concatenated_in_order():
    major, pts, feats = hand_case()
    a = pool_features_ref(major, [(pts, feats)], np.zeros(3), radius=R)
    b = pool_features_ref(major, [(pts[:3], feats[:3]), (pts[3:], feats[3:])], np.zeros(3), radius=R)
    assert np.array_equal(a, b)
    assert not pool_features_ref(major, [], np.zeros(3)).any()


def test_host_side_transform_and_fail_loudly_without_gpu():
    """Host logic of autoinst_b200.pooling: the rigid transform of `transform_pcd` (point_cloud_utils.py:24-35) and no CPU
    fallback for the pooling itself."""
    import torch
    from autoinst_b200.pooling import transform_points, tarl_features_per_patch
    rng = np.random.default_rng(0)
    a = 0.7
    T = np.eye(4)
    T[:3, :3] = [[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]]
    T[:3, 3] = [1.0, -2.0, 0.5]
    pts = rng.normal(size=(50, 3))
    hom = (T @ np.concatenate([pts, np.ones((50, 1))], axis=1).T).T
    assert np.allclose(transform_points(pts, T), hom[:, :3], rtol=0, atol=1e-14)
    # raw KITTI scans are n x 4 (x, y, z, remission): the reference slices points[:, :3] (point_cloud_utils.py:24-35)
    pts4 = np.concatenate([pts, rng.random((50, 1))], axis=1)
    assert np.array_equal(transform_points(pts4, T), transform_points(pts, T))
    if torch.cuda.is_available():
        return

    class Dataset:
        def get_pose(self, i): return np.eye(4)
        def get_point_cloud(self, i): return pts
        def get_tarl_features(self, i): return np.ones((50, 96), dtype=np.float32)
    with pytest.raises(Exception):                      # the CUDA library cannot create a handle: nothing is computed on the CPU
        tarl_features_per_patch(Dataset(), pts, np.eye(4), np.zeros(3), [0])


@pytest.mark.gpu
def test_pool_hand_case_gpu(cuda_device):
    from autoinst_b200 import api
    major, pts, feats = hand_case()
    out, cnt = api.feature_pool(major, pts, feats, R, -12.5 * np.ones(3), 12.5 * np.ones(3), return_count=True,
                                device=cuda_device)
    ref, rc = pool_features_ref(major, [(pts, feats)], np.zeros(3), radius=R, return_count=True)
    assert cnt.cpu().numpy().tolist() == rc.tolist() == [2, 1, 0]
    assert np.array_equal(out.cpu().numpy(), ref)              # two terms: no rounding freedom


@pytest.mark.gpu
def test_pool_matches_the_reference_golden(cuda_device):
    from autoinst_b200 import api
    g, scans = load_pooling_golden()
    half = 0.5 * g["chunk_size"]
    out, cnt = api.feature_pool(g["major"], g["scan_points"], g["scan_features"], float(g["radius"]), g["center"] - half,
                                g["center"] + half, return_count=True, device=cuda_device)
    assert np.array_equal(cnt.cpu().numpy(), g["count"])       # the reference's neighbour sets, point for point
    out = out.cpu().numpy()
    assert np.array_equal(out[g["count"] == 0], g["out"][g["count"] == 0])
    assert np.allclose(out, g["out"], rtol=1e-13, atol=1e-14)  # float64 means, different summation order


@pytest.mark.gpu
@pytest.mark.parametrize("fdim,normalise", [(96, False), (96, True), (384, False), (7, False)])
def test_pool_matches_oracle_on_a_chunk(cuda_device, fdim, normalise):
    from autoinst_b200 import api
    ch = make_chunk(11, n_target=2500, features="tarl", center=(40.0, -7.0, 1.5))
    scans = make_scans(ch, n_scans=6, pts_per_major=2.5, fdim=fdim)
    ref, rc = pool_features_ref(ch.points, scans, ch.center, radius=R, normalise=normalise, return_count=True)
    pts = np.concatenate([s[0] for s in scans])
    fts = np.concatenate([s[1] for s in scans])
    out, cnt = api.feature_pool(ch.points, pts, fts, R, ch.center - 12.5, ch.center + 12.5, normalise=normalise,
                                return_count=True, device=cuda_device)
    out, cnt = out.cpu().numpy(), cnt.cpu().numpy()
    assert np.array_equal(cnt, rc)                             # same neighbour sets, point for point
    assert (rc == 0).any() and (rc > 8).any()
    assert np.array_equal(out[rc == 0], ref[rc == 0])          # zero rows stay exactly zero
    # float64 means of <= 60 float32 values in a different summation order: rounding level
    assert np.allclose(out, ref, rtol=1e-13, atol=1e-14)
    again = api.feature_pool(ch.points, pts, fts, R, ch.center - 12.5, ch.center + 12.5, normalise=normalise,
                             device=cuda_device).cpu().numpy()
    assert np.array_equal(again, out)                          # deterministic


@pytest.mark.gpu
def test_pool_edge_cases(cuda_device):
    from autoinst_b200 import api
    major = np.array([[0.0, 0.0, 0.0], [1.0, 1.0, 1.0]])
    lo, hi = -12.5 * np.ones(3), 12.5 * np.ones(3)
    out = api.feature_pool(major, np.zeros((0, 3)), np.zeros((0, 96), dtype=np.float32), R, lo, hi, device=cuda_device)
    assert out.shape == (2, 96) and not out.cpu().numpy().any()                 # no scan point at all
    pts = np.array([[20.0, 0.0, 0.0], [0.05, 0.0, 30.0]])                       # every point outside the cube
    out, cnt = api.feature_pool(major, pts, np.ones((2, 96), dtype=np.float32), R, lo, hi, return_count=True,
                                device=cuda_device)
    assert not out.cpu().numpy().any() and not cnt.cpu().numpy().any()
    # a major point outside the cube still collects the points inside the cube that are within the radius
    major2 = np.array([[12.55, 0.0, 0.0]])
    pts2 = np.array([[12.45, 0.0, 0.0], [12.6, 0.0, 0.0]])
    f2 = np.array([[1.0, 2.0], [10.0, 20.0]], dtype=np.float32)
    out, cnt = api.feature_pool(major2, pts2, f2, R, lo, hi, return_count=True, device=cuda_device)
    assert cnt.cpu().numpy().tolist() == [1] and out.cpu().numpy().tolist() == [[1.0, 2.0]]
    with pytest.raises(Exception):
        api.feature_pool(major, pts, np.ones((2, 96), dtype=np.float32), 0.0, lo, hi, device=cuda_device)


@pytest.mark.gpu
def test_tarl_features_per_patch_drop_in(cuda_device):
    """Reference signature (`chunk_generation.py:205-211`) with a fake dataset: poses move every scan into its own
    lidar frame, the function has to bring it back (`:228-231`)."""
    from autoinst_b200.pooling import tarl_features_per_patch
    ch = make_chunk(12, n_target=1500, features="tarl", center=(3.0, 4.0, 0.5))
    scans = make_scans(ch, n_scans=4, pts_per_major=2.0)
    rng = np.random.default_rng(4)

    def pose(k):
        a = 0.3 * k
        T = np.eye(4)
        T[:3, :3] = [[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]]
        T[:3, 3] = [2.0 * k, -1.0 * k, 0.1 * k]
        return T
    T_pcd = pose(7)

    class FakeDataset:
        def get_pose(self, i): return pose(i)
        def get_tarl_features(self, i): return scans[i][1]
        def get_point_cloud(self, i):                        # chunk frame -> lidar frame of scan i
            T = np.linalg.inv(np.linalg.inv(T_pcd) @ pose(i))
            return scans[i][0] @ T[:3, :3].T + T[:3, 3]

    class Pcd:
        points = ch.points
    got = tarl_features_per_patch(FakeDataset(), Pcd(), T_pcd, ch.center, [0, 1, 2, 3])
    ref, rc = pool_features_ref(ch.points, scans, ch.center, radius=R, return_count=True)
    assert got.shape == (ch.n, 96) and got.dtype == np.float64
    # the round trip through the poses moves points by ~1e-15 m: a neighbour can only change where a point sits
    # within that distance of the sphere or of a cube face; everything else agrees to rounding
    same = np.isclose(got, ref, rtol=1e-9, atol=1e-12).all(axis=1)
    assert same.mean() > 0.999
    assert np.array_equal(~got.any(axis=1), rc == 0) or same.mean() > 0.999
