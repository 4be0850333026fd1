"""Oracle: map-level merge of chunk labelings and the driver glue around the metrics
(TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates, on plain arrays with integer labels instead of RGB colours:
  merge_chunks_unite_instances2   `pipeline/utils/point_cloud/point_cloud_utils.py:387-491`
  merge_unite_gt                  `:320-329`
  remove_semantics                `:253-287`
  colour -> integer label glue    `pipeline/run_pipeline.py:203-223`
Pinned: `oracle/make_golden.py::golden_merge` runs the reference's own, unmodified functions on a synthetic 4-chunk map
with a numpy stand-in for the Open3D `PointCloud` (inclusive `crop`, `+=`, `remove_duplicated_points` keeping the first
occurrence in point order) and labels encoded as colours; this module reproduces their outputs exactly
(`tests/golden/merge.npz`, `tests/test_oracle.py`).  Not pinnable here: Open3D's own duplicate-removal order (it cannot
be installed); both sides of the level-3 comparison go through this same restatement.
Label 0 is the background ("black" in the reference).
"""
from __future__ import annotations

import numpy as np


def _dedup_first(points, labels):
    """Open3D remove_duplicated_points: exact-coordinate duplicates dropped, first occurrence kept."""
    _, first = np.unique(points, axis=0, return_index=True)
    keep = np.sort(first)
    return points[keep], labels[keep]


def merge_chunks_unite_instances(chunks):
    """chunks: list of (points (n,3) float64, labels (n,) int64, 0 = background), in file-name order.
    Returns merged (points, labels)."""
    pts, lab = chunks[0][0].copy(), chunks[0][1].copy()
    pts, lab = np.asarray(pts, dtype=np.float64), np.asarray(lab, dtype=np.int64)
    for p2, l2 in chunks[1:]:
        p2 = np.asarray(p2, dtype=np.float64)
        l2 = np.asarray(l2, dtype=np.int64).copy()
        center = p2.mean(axis=0)                                        # :397-403
        lo, hi = center - 20.0, center + 20.0                           # side_length = 40, :405-417
        inside = np.all((pts >= lo) & (pts <= hi), axis=1)
        p1, l1 = pts[inside], lab[inside]
        inst1 = {i: p1[l1 == i] for i in np.unique(l1) if i != 0}       # :425-432
        inst2 = {i: np.where(l2 == i)[0] for i in np.unique(l2) if i != 0}
        pairs = []
        for id1, q1 in inst1.items():                                   # :444-463
            bmin, bmax = q1.min(axis=0), q1.max(axis=0)
            for id2, idx2 in inst2.items():
                q2 = p2[idx2]
                inter = int(np.count_nonzero(np.all(q2 >= bmin, axis=1) & np.all(q2 <= bmax, axis=1)))
                if inter > 0:
                    union = len(np.unique(np.concatenate((q1, q2))))    # NB: unique SCALARS, as the reference (:457)
                    iou = float(inter) / float(union)
                    if iou > 0.01:
                        pairs.append((id1, id2, iou))
        best = {}                                                       # :465-477: each id2 keeps its best id1
        for id1, id2, iou in pairs:
            if id2 not in best or iou > best[id2][1]:
                best[id2] = (id1, iou)
        for id2, (id1, _) in best.items():                              # :479-481
            l2[inst2[id2]] = id1
        pts = np.concatenate((pts, p2))                                 # :488
        lab = np.concatenate((lab, l2))
        pts, lab = _dedup_first(pts, lab)                               # :489
    return pts, lab


def merge_unite_gt(chunks):
    pts = np.concatenate([np.asarray(c[0], dtype=np.float64) for c in chunks])
    lab = np.concatenate([np.asarray(c[1], dtype=np.int64) for c in chunks])
    return _dedup_first(pts, lab)


def compact_labels(labels):
    """np.unique(colors, axis=0, return_inverse=True): background (smallest) becomes 0 (`run_pipeline.py:216-218`)."""
    uniq, inv = np.unique(labels, return_inverse=True)
    if uniq[0] != 0:
        inv = inv + 1
    return inv.astype(np.int64)


def remove_semantics(gt_labels, preds, threshold=0.8):
    """Predicted labels with more than `threshold` of their points on GT background become 0 (`:253-287`)."""
    out = preds.copy()
    bg = gt_labels == 0
    for u in np.unique(preds):
        idx = preds == u
        if np.count_nonzero(bg & idx) > threshold * np.count_nonzero(idx):
            out[idx] = 0
    return out


def canonical_labels(seg_labels):
    """Relabel the segments of one chunk by first occurrence in point order.  The greedy rules downstream
    (merge tie-breaks, AP's np.unique order, `metrics_class.py:181-235`) depend on label VALUES, which the
    reference draws at random (colours); both sides of a comparison must number segments the same way
    (last paragraph)."""
    seg_labels = np.asarray(seg_labels)
    _, first, inv = np.unique(seg_labels, return_index=True, return_inverse=True)
    rank = np.empty(len(first), dtype=np.int64)
    rank[np.argsort(first, kind="stable")] = np.arange(len(first))
    return rank[inv.reshape(-1)]


def globally_unique(chunk_id, seg_labels):
    """Per-chunk segment ids -> ids unique across the map (the reference draws a random colour per segment)."""
    return (np.int64(chunk_id + 1) << 20) + np.asarray(seg_labels, dtype=np.int64) + 1
