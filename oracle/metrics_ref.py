"""Oracle: instance metrics (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates `pipeline/metrics/metrics_class.py` (Metrics.update_stats :137-179, filter_labels :302-309,
get_tp_fp / calculate_full_stats :61-117,315-340, average_precision :181-235) and
`pipeline/metrics/modified_LSTQ.py` (evaluator.add_batch :23-55, get_eval :57-80) for ONE update_stats call
(one map), which is what `run_pipeline.py:234-243` does per sequence with TEST_MAP.
Pinned: `oracle/make_golden.py` runs the reference's own Metrics class (imported with a matplotlib stub) on
seeded label arrays and stores its outputs in tests/golden/metrics.npz; tests/test_oracle.py replays them.
IoU tables come from one pass over (pred, gt) pairs instead of `np.intersect1d` per pair — same values.
"""
from __future__ import annotations

import numpy as np

OVERLAPS = [0.25, 0.5, 0.55, 0.6, 0.65, 0.7, 0.75, 0.8, 0.85, 0.9, 0.95]      # metrics_class.py:40
AP_OVERLAPS = OVERLAPS[1:]                                                     # :41


def filter_labels(label, min_points):
    label = label.copy()
    ids, cnt = np.unique(label, return_counts=True)
    for i, c in zip(ids, cnt):
        if c < min_points:
            label[label == i] = 0                                              # :302-309
    return label


def _iou_table(pred, gt):
    """iou[(p, g)] for every co-occurring pair of non-zero labels; sizes of every label."""
    psz = dict(zip(*np.unique(pred, return_counts=True)))
    gsz = dict(zip(*np.unique(gt, return_counts=True)))
    both = (pred != 0) & (gt != 0)
    key = np.stack([pred[both], gt[both]], axis=1)
    table = {}
    if key.shape[0]:
        uk, cnt = np.unique(key, axis=0, return_counts=True)
        for (p, g), c in zip(uk, cnt):
            table[(int(p), int(g))] = c / float(psz[p] + gsz[g] - c)          # Metrics.iou :296-300
    return table, psz, gsz


def _greedy_matches(pred_ids, gt_ids, table, thr):
    """First unused GT (in np.unique order) with IoU >= thr, per prediction in order (:77-96, :213-224)."""
    used = set()
    out = []
    for p in pred_ids:
        hit = None
        for g in gt_ids:
            if table.get((p, g), 0.0) >= thr and g not in used:
                hit = g
                used.add(g)
                break
        out.append(hit)
    return out


def instance_metrics(all_labels, pred_labels, gt_labels, min_points=200):
    """Metrics(...).update_stats(all_labels, pred_labels, gt_labels) for a fresh Metrics object.
    Returns dict with p, r, f1, ap, ap0.25, ap0.5, S_assoc (the keys of sequence_stats, :262-269)."""
    pred = filter_labels(np.asarray(pred_labels), min_points)                  # :147
    allp = filter_labels(np.asarray(all_labels), min_points)                   # :148
    gt = np.asarray(gt_labels)
    table, psz, gsz = _iou_table(pred, gt)
    pred_ids = [int(x) for x in np.unique(pred) if x != 0]
    gt_ids = [int(x) for x in np.unique(gt) if x != 0]

    m = _greedy_matches(pred_ids, gt_ids, table, 0.5)                          # calculate_full_stats
    tps = sum(h is not None for h in m)
    n_gt = (np.unique(gt).shape[0] - 1) if 0 in gt else 0                      # :321-322
    n_pred = np.unique(pred).shape[0] - 1                                      # :323
    prec = tps / n_pred
    rec = tps / n_gt
    try:
        f1 = 2 * (prec * rec) / (prec + rec)
    except ZeroDivisionError:
        f1 = 0

    aps = {}
    for thr in OVERLAPS:                                                       # average_precision :181-235
        precision, recall = [1.0], [0.0]
        tp = fp = 0
        fn = len(gt_ids)
        for h in _greedy_matches(pred_ids, gt_ids, table, thr):
            if h is not None:
                tp += 1
                fn -= 1
            else:
                fp += 1
            precision.append(tp / float(tp + fp))
            recall.append(tp / float(tp + fn))
        trap = getattr(np, "trapezoid", None) or np.trapz                         # np.trapz in the reference (:234)
        aps[thr] = float(trap(precision, recall))
    ap = sum(aps[o] for o in AP_OVERLAPS) / float(len(AP_OVERLAPS))

    s_assoc = lstq_association(allp, gt, min_points)
    return {"p": prec, "r": rec, "f1": f1, "ap": ap, "ap0.25": aps[0.25], "ap0.5": aps[0.5], "S_assoc": s_assoc}


def lstq_association(pred_labels, gt_labels, min_points=200):
    """evaluator.add_batch + get_eval for one batch (modified_LSTQ.py:23-80)."""
    pred_labels = np.asarray(pred_labels)
    gt_labels = np.asarray(gt_labels)
    vp = pred_labels[(pred_labels != 0) & (pred_labels != -1)]
    vg = gt_labels[gt_labels != 0]
    pl, pa = np.unique(vp, return_counts=True)
    gl, ga = np.unique(vg, return_counts=True)
    keep = ga > min_points                                                     # strict '>' (:31-32)
    gl, ga = gl[keep], ga[keep]
    both = (pred_labels > 0) & (gt_labels > 0)
    inter = {}
    if both.any():
        uk, cnt = np.unique(np.stack([pred_labels[both], gt_labels[both]], axis=1), axis=0, return_counts=True)
        inter = {(int(p), int(g)): int(c) for (p, g), c in zip(uk, cnt)}
    parea = {int(l): int(a) for l, a in zip(pl, pa)}
    outer = 0.0
    for g, garea in zip(gl, ga):
        inner = 0.0
        for p, area in parea.items():
            t = inter.get((p, int(g)))
            if t is not None:
                inner += t * (t / (garea + area - t))
        outer += float(inner) / float(garea)
    if len(gl) == 0:
        return float("nan")
    return outer / len(gl)
