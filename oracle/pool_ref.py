"""Oracle: TARL feature pooling onto the major voxel points (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates `tarl_features_per_patch`, `pipeline/utils/point_cloud/chunk_generation.py:205-258`, on plain arrays:
  crop of every scan to the chunk cube, strict comparisons        :220-221, 233-236
  concatenation in scan order                                      :238-242
  radius search with MAJOR_VOXEL_SIZE / 2 around every major point :212, 249-250
  `np.mean(features_in_radius, axis=0)`, zero row when empty       :251-256
  optional L2 normalisation (TARL_NORM, False in config.py:64)     :253-254

Third-party arithmetic: the radius search is Open3D 0.17 `KDTreeFlann.search_radius_vector_3d` (pinned in
`setup.sh:9`, not vendored, not installable here).  Its published algorithm: nanoflann `radiusSearch` with
`radius * radius`, whose `RadiusResultSet::addPoint` keeps a point iff `dist < radius` on SQUARED distances, results
sorted by distance.  Restated with scipy's cKDTree (candidates within the radius, then the strict float64 test and
the distance sort).
Pinned: `oracle/make_golden.py::golden_pooling` runs the reference's own, unmodified `tarl_features_per_patch` on
synthetic scans (with that cKDTree restatement standing in for Open3D's KD-tree and a numpy `PointCloud.transform`)
and this module reproduces its output bit for bit; inputs and output are committed as `tests/golden/pooling.npz`.
What no fixture here can pin is Open3D's own behaviour exactly on the sphere and its result order (the latter moves
a mean by rounding only).
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import cKDTree


def crop_scan(coords, feats, center_position, chunk_size=(25.0, 25.0, 25.0)):
    """:220-221, 233-238 — strict box test around the chunk centre."""
    center = np.asarray(center_position, dtype=np.float64)
    half = 0.5 * np.asarray(chunk_size, dtype=np.float64)
    lo, hi = center - half, center + half
    keep = np.where(np.all(coords > lo, axis=1) & np.all(coords < hi, axis=1))[0]
    return coords[keep], feats[keep]


def pool_features_ref(major_points, scans, center_position, *, radius=0.35 / 2.0, chunk_size=(25.0, 25.0, 25.0),
                      normalise=False, return_count=False):
    """scans: list of (coords (m,3) float64 in the chunk frame, feats (m,F)) in scan order.
    Returns (n_major, F) float64."""
    major_points = np.asarray(major_points, dtype=np.float64)
    fdim = scans[0][1].shape[1] if scans else 96
    cat_p = np.zeros((0, 3))
    cat_f = np.zeros((0, fdim))
    for coords, feats in scans:
        c, f = crop_scan(np.asarray(coords, dtype=np.float64), np.asarray(feats), center_position, chunk_size)
        cat_p = np.concatenate((cat_p, c))
        cat_f = np.concatenate((cat_f, f))                      # float64 from here on, as in the reference
    out = np.zeros((major_points.shape[0], fdim))
    cnt = np.zeros(major_points.shape[0], dtype=np.int32)
    if cat_p.shape[0] == 0:
        return (out, cnt) if return_count else out
    tree = cKDTree(cat_p)
    r2 = radius * radius
    for i, p in enumerate(major_points):
        cand = np.asarray(tree.query_ball_point(p, radius * (1.0 + 1e-9)), dtype=np.int64)
        if cand.size == 0:
            continue
        d2 = ((cat_p[cand] - p) ** 2).sum(axis=1)
        keep = d2 < r2                                          # nanoflann: strictly inside
        idx = cand[keep][np.argsort(d2[keep], kind="stable")]   # sorted by distance
        if idx.size == 0:
            continue
        out[i, :] = np.mean(cat_f[idx], axis=0)
        cnt[i] = idx.size
        if normalise:
            out[i] /= np.linalg.norm(out[i])
    return (out, cnt) if return_count else out
