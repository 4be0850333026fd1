"""Oracle: DINOv2 feature look-up per view and the mean over views (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates, on plain arrays, the DINOv2 half of `image_based_features_per_patch`
(`pipeline/utils/image/image_utils.py:91-352`, the part behind the visibility bookkeeping) and `dinov2_mean`
(`:363-371`), plus `point_to_pixel` (`pipeline/utils/image/point_to_pixels.py:6-38`):
  per view  major points into the camera frame (Open3D `transform`)                      :264
            1-NN of every major point among the view's visible chunk points, kept if the distance is
            strictly below MAJOR_VOXEL_SIZE / 2                                            :266-276
            projection K p, division by depth, np.round, image bounds, depth > 0          point_to_pixels.py:21-30
            feature-map pixel = int(factor * pixel), factor = map size / image size       :255-256, 341-346
  mean      over the views whose looked-up feature vector has any non-zero entry          :363-371
What stays on the reference's side (and outside this oracle): poses and calibration, hidden point removal
(Open3D convex hull) or the `hpr_masks` argument, the statistical-outlier filter, the set intersections (:142-212).
Third-party arithmetic: Open3D's KD-tree (nearest neighbour: any implementation returns the same minimum distance),
`PointCloud.transform` (homogeneous 4 x 4 product, restated with numpy as in oracle/make_golden.py).
Pinned: `oracle/make_golden.py::golden_dino` runs the reference's own, unmodified `image_based_features_per_patch`
and `dinov2_mean` on a synthetic scene (stand-ins for the Open3D pieces, `hpr_masks` given) and this module
reproduces the result bit for bit (`tests/golden/dino.npz`).
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import cKDTree


def transform_points(points, T):
    """Open3D PointCloud.transform: homogeneous product, division by w."""
    P = np.asarray(points, dtype=np.float64)
    hom = np.concatenate([P, np.ones((P.shape[0], 1))], axis=1) @ np.asarray(T, dtype=np.float64).T
    return hom[:, :3] / hom[:, 3:4]


def view_pixels_ref(major_cam, visible_cam, K, img_h, img_w, map_h, map_w, max_dist):
    """Feature-map pixel (row, col) of every major point in one view, or (-1, -1).  Returns int64 [N, 2]."""
    major_cam = np.asarray(major_cam, dtype=np.float64)
    out = np.full((major_cam.shape[0], 2), -1, dtype=np.int64)
    visible_cam = np.asarray(visible_cam, dtype=np.float64)
    if visible_cam.shape[0] == 0:
        return out
    tree = cKDTree(visible_cam)
    _, nn = tree.query(major_cam, k=1)
    nc_indices = [j for j, point in enumerate(major_cam)
                  if np.linalg.norm(point - visible_cam[nn[j]]) < max_dist]              # :271-276, strict
    if not nc_indices:
        return out
    pts = major_cam[nc_indices]
    img = np.asarray(K, dtype=np.float64) @ pts.transpose()                             # point_to_pixels.py:21
    img[:2, :] /= img[2, :]
    img[:2, :] = np.round(img[:2, :])
    inds = np.where((img[0, :] < img_w) & (img[0, :] >= 0) & (img[1, :] < img_h) & (img[1, :] >= 0) & (img[2, :] > 0))[0]
    f0 = map_h / img_h                                                                   # :255-256
    f1 = map_w / img_w
    for ind in inds:
        pixel = img[:2, ind].astype(int)
        out[nc_indices[ind], 0] = int(f0 * pixel[1])                                     # :341-342
        out[nc_indices[ind], 1] = int(f1 * pixel[0])
    return out


def dino_mean_ref(major_points, views, max_dist=0.35 / 2.0, fdim=384):
    """views: list of dicts with T_pcd2cam (4x4), visible_cam (M,3) float64 (the view's visible chunk points in the camera
    frame), K (3x3), img_hw (h, w), feature_map (Hp, Wp, F) float32 — or None for a skipped view (:183-191, :212-214).
    Returns the (N, F) float64 array `dinov2_mean(point2dino)` gives for this camera."""
    major_points = np.asarray(major_points, dtype=np.float64)
    n = major_points.shape[0]
    feats = [[] for _ in range(n)]
    for view in views:
        if view is None:
            continue
        fmap = np.asarray(view["feature_map"])
        h, w = view["img_hw"]
        pix = view_pixels_ref(transform_points(major_points, view["T_pcd2cam"]), view["visible_cam"], view["K"], h, w,
                              fmap.shape[0], fmap.shape[1], max_dist)
        for j in np.where(pix[:, 0] >= 0)[0]:
            f = fmap[pix[j, 0], pix[j, 1], :].astype(np.float64)
            if f.any():                                                                  # non_zero_mask, :366
                feats[j].append(f)
    out = np.zeros((n, fdim))
    for j in range(n):
        if feats[j]:
            out[j] = np.mean(np.stack(feats[j]), axis=0)                                 # :369-370
    return out
