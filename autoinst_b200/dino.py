"""DINOv2 features of the major voxel points — "next" row N3 of SURVEY.md §8f, DINOv2 half.

`dinov2_mean_per_patch` does what `ncuts_chunk` gets from `image_based_features_per_patch(..., sam=False, dino=True)`
followed by `dinov2_mean` per camera (`pipeline/ncuts/ncuts_utils.py:81-110`, `pipeline/utils/image/image_utils.py:
91-371`).  The host part keeps the reference's own steps and helpers (poses and calibration, the statistical-outlier
filter of the chunk, hidden point removal or the `hpr_masks` argument, the index-set intersection, `:105-214`); the
per-view nearest-neighbour test of every major point, the projection, the feature-map look-up and the mean over views
— Python loops over points and views with an Open3D KD-tree and an N x views x 384 float64 array in the reference
(`:264-352`, `:363-371`) — run on the GPU through the C ABI (`ancuts_dino_view_pixels`, `ancuts_dino_mean`).  There is
no CPU fallback.
"""
import copy
import warnings

import numpy as np

try:                                            # cwd = pipeline/ in the reference layout
    from config import CAM_IDS, HPR_RADIUS, MAJOR_VOXEL_SIZE, NUM_DINO_FEATURES      # config.py:56,66,67,72
except ModuleNotFoundError as _e:               # stand-alone use ONLY; any other failure of config.py propagates
    if _e.name != "config":
        raise
    warnings.warn("autoinst_b200.dino: no `config` module on sys.path; using CAM_IDS [0], HPR_RADIUS 1000, "
                  "MAJOR_VOXEL_SIZE 0.35, NUM_DINO_FEATURES 384", RuntimeWarning)
    CAM_IDS, HPR_RADIUS, MAJOR_VOXEL_SIZE, NUM_DINO_FEATURES = [0], 1000, 0.35, 384


def _transform(points, T):
    """Open3D PointCloud.transform on an array: homogeneous product, division by w."""
    P = np.asarray(points, dtype=np.float64)
    hom = np.concatenate([P, np.ones((P.shape[0], 1))], axis=1) @ np.asarray(T, dtype=np.float64).T
    return hom[:, :3] / hom[:, 3:4]


def build_views(dataset, pcd, chunk_indices, T_pcd2world, cam_indices, cam_name, hpr_masks=None):
    """The reference's per-view bookkeeping (image_utils.py:105-214) up to the list of visible chunk points per view.
    Returns a list with one entry per cam index: None for a skipped view, else the dict `api.dino_mean_views` takes."""
    from utils.point_cloud.point_cloud_utils import get_statistical_inlier_indices, get_subpcd
    pcd_pts = np.asarray(pcd.points)
    chunk_indices = np.asarray(chunk_indices)
    pcd_chunk = get_subpcd(pcd, chunk_indices)                                   # :105
    inliers = get_statistical_inlier_indices(pcd_chunk)                          # :106
    chunk_and_inlier = set(chunk_indices[np.asarray(inliers)].tolist())          # :107
    w, h = dataset.get_image(cam_name, 0).size                                   # :136-138
    if hpr_masks is not None:
        assert len(cam_indices) == hpr_masks.shape[0]                            # :142-143
    pts = np.asarray(pcd_chunk.points)
    min_bound, max_bound = pts.min(axis=0), pts.max(axis=0)                      # :162-166
    views = []
    for i, points_index in enumerate(cam_indices):
        T_world2lidar = np.linalg.inv(dataset.get_pose(points_index))            # :146-147
        T_lidar2cam, K = dataset.get_calibration_matrices(cam_name)
        T_pcd2cam = T_lidar2cam @ T_world2lidar @ T_pcd2world                    # :149-150
        cam_pts = _transform(pcd_pts, T_pcd2cam)                                 # :157
        if hpr_masks is None:
            from utils.image.hidden_points_removal import hidden_point_removal_o3d
            world = _transform(pcd_pts, dataset.get_pose(0))                     # :156
            bound = np.where(np.all(world > min_bound, axis=1) & np.all(world < max_bound, axis=1))[0]     # :171-180
            try:
                vis = hidden_point_removal_o3d(cam_pts[bound], camera=[0, 0, 0], radius_factor=HPR_RADIUS)   # :184-189
            except Exception:
                print("hpr skip")                                                # :190-192
                views.append(None)
                continue
            visible = bound[np.asarray(vis, dtype=np.int64)]                     # :204
        else:
            visible = np.where(hpr_masks[i])[0]                                  # :207
        frame = list(set(visible.tolist()) & chunk_and_inlier)                   # :209
        if len(frame) == 0:
            print("out of view skip")                                            # :213-215
            views.append(None)
            continue
        views.append(dict(T_pcd2cam=T_pcd2cam, visible_cam=cam_pts[frame], K=np.asarray(K, dtype=np.float64),
                          img_hw=(h, w), feature_map=dataset.get_dinov2_features(cam_name, points_index)))   # :243-246
    return views


def dinov2_mean_per_patch(dataset, pcd, chunk_indices, chunk_nc, T_pcd2world, cam_indices, hpr_masks=None, device=None):
    """[dinov2_mean(p) for p in image_based_features_per_patch(..., sam=False, dino=True)[0]]: one (num_points, 384)
    float64 array per camera of CAM_IDS, zero rows for points no view sees."""
    from autoinst_b200 import api
    if NUM_DINO_FEATURES != 384:
        raise NotImplementedError("NUM_DINO_FEATURES < 384 (UMAP reduction, image_utils.py:224-241) is not supported")
    cams = ["cam2", "cam3"]                                                      # :104
    major = np.asarray(copy.deepcopy(chunk_nc).points) if hasattr(chunk_nc, "points") else np.asarray(chunk_nc)
    out = []
    for cam_id in CAM_IDS:
        views = build_views(dataset, pcd, chunk_indices, T_pcd2world, cam_indices, cams[cam_id], hpr_masks)
        out.append(api.dino_mean_views(major, views, MAJOR_VOXEL_SIZE / 2, feat_dim=NUM_DINO_FEATURES, device=device).cpu().numpy())
    return out
