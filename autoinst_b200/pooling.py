"""TARL feature pooling onto the major voxel points — "next" row N3 of SURVEY.md §8f.

`tarl_features_per_patch` has the reference's signature and argument meaning
(`pipeline/utils/point_cloud/chunk_generation.py:205-258`; called by `ncuts_chunk`, `ncuts_utils.py:135-141`).
The per-scan part stays what the reference does on the host (fetch features and points from the dataset,
move them into the chunk frame, `:223-231`); the crop to the chunk cube, the radius search and the means
— a KD-tree query per major point in a Python loop in the reference (`:244-256`) — run on the GPU through
the C ABI `ancuts_feature_pool`.  There is no CPU fallback.
"""
import numpy as np

import warnings

try:                                            # cwd = pipeline/ in the reference layout
    from config import MAJOR_VOXEL_SIZE, CHUNK_SIZE, TARL_NORM      # config.py:56,57,64
except ModuleNotFoundError as _e:               # stand-alone use ONLY: the shipped values; any other failure of config.py
    if _e.name != "config":                     # propagates (a wrong cwd must not silently change the parameters)
        raise
    warnings.warn("autoinst_b200.pooling: no `config` module on sys.path; using MAJOR_VOXEL_SIZE 0.35, CHUNK_SIZE 25, "
                  "TARL_NORM False", RuntimeWarning)
    MAJOR_VOXEL_SIZE = 0.35
    CHUNK_SIZE = np.array([25, 25, 25])
    TARL_NORM = False


def transform_points(coords, T):
    """`transform_pcd` (point_cloud_utils.py:24-35, Open3D `transform`): rigid 4 x 4 transform of n x 3 points."""
    coords = np.asarray(coords, dtype=np.float64)[:, :3]        # raw KITTI scans are n x 4 (get_pcd(points[:, :3]), :24-35)
    T = np.asarray(T, dtype=np.float64)
    return coords @ T[:3, :3].T + T[:3, 3]


def pool_scan_features(major_points, scan_points, scan_features, center_position, *, radius=None, chunk_size=None,
                       normalise=None, device=None, return_count=False):
    """Array-level core: concatenated scan points (chunk frame) + features -> n_major x F float64 numpy array."""
    from autoinst_b200 import api
    chunk_size = np.asarray(CHUNK_SIZE if chunk_size is None else chunk_size, dtype=np.float64)
    center = np.asarray(center_position, dtype=np.float64)
    lo, hi = center - 0.5 * chunk_size, center + 0.5 * chunk_size          # :220-221
    r = MAJOR_VOXEL_SIZE / 2. if radius is None else radius                # :212
    out = api.feature_pool(major_points, scan_points, scan_features, r, lo, hi,
                           normalise=TARL_NORM if normalise is None else normalise, return_count=return_count,
                           device=device)
    if return_count:
        return out[0].cpu().numpy(), out[1].cpu().numpy()
    return out.cpu().numpy()


def tarl_features_per_patch(dataset, pcd, T_pcd, center_position, tarl_indices):
    """Same contract as the reference: (num_points, 96) float64, zero rows where no scan point is in range."""
    inv_T_pcd = np.linalg.inv(T_pcd)
    pts, feats = [], []
    for points_index in tarl_indices:
        f = np.asarray(dataset.get_tarl_features(points_index))            # :226
        c = np.asarray(dataset.get_point_cloud(points_index))              # :227
        T_local2global = inv_T_pcd @ dataset.get_pose(points_index)        # :230-231
        pts.append(transform_points(c, T_local2global))
        feats.append(f.astype(np.float32, copy=False))
    major = np.asarray(pcd.points) if hasattr(pcd, "points") else np.asarray(pcd)
    if not pts:
        return np.zeros((major.shape[0], 96))
    return pool_scan_features(major, np.concatenate(pts), np.concatenate(feats), center_position)
