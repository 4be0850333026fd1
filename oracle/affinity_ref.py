"""Oracle: dense pairwise affinity of one chunk (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates `pipeline/ncuts/ncuts_utils.py:56-67` (spatial term), `:112-113,125-133` (DINOv2 term),
`:135-149` (TARL term), `:151-156` (product) and `pipeline/utils/point_cloud/point_cloud_utils.py:189-195`
(isolated rows) of the reference on plain arrays, in float64 with `scipy.spatial.distance.cdist`
exactly as the reference does.  The SAM term (`ncuts_utils.py:115-123`) is the identity because
beta = 0.0 in every shipped config (`config.py:12,23,34,45`).
"""
from __future__ import annotations

import numpy as np
from scipy.spatial.distance import cdist


def affinity_ref(points, tarl=None, dino=None, *, alpha=1.0, theta=0.0, gamma=0.0,
                 proximity=1.0):
    """Return the N×N float64 affinity A the reference hands to `normalized_cut`.

    points: (N,3); tarl: (N,96) or None; dino: (N,384) or a list of per-camera arrays or None.
    A falsy gain switches its term off exactly as the reference's `if CONFIG[...]` does.
    """
    points = np.asarray(points, dtype=np.float64)
    sd = cdist(points, points)                                  # ncuts_utils.py:60
    mask = np.where(sd <= proximity, 1, 0)                      # :61  (inclusive)
    spatial = mask * np.exp(-alpha * sd) if alpha else mask     # :63-66

    dino_w = mask.copy()                                        # :113
    if gamma:
        if dino is None or (isinstance(dino, (list, tuple)) and len(dino) == 0):
            raise ValueError("The length should be longer than 0!")       # :126-127
        cams = dino if isinstance(dino, (list, tuple)) else [dino]
        for g in cams:                                          # :129-133
            g = np.asarray(g, dtype=np.float64)
            dino_w = dino_w * np.exp(-gamma * cdist(g, g))

    if theta:
        t = np.asarray(tarl, dtype=np.float64)
        none = ~t.any(axis=1)                                   # :143
        td = cdist(t, t)                                        # :144
        td[none] = 0                                            # :145
        td[:, none] = 0                                         # :146
        tarl_w = mask * np.exp(-theta * td)                     # :147
    else:
        tarl_w = mask                                           # :149

    sam_w = mask                                                # :112 (beta == 0)
    return tarl_w * spatial * sam_w * dino_w                    # :151-156


def drop_isolated(A):
    """`remove_isolated_points` (`point_cloud_utils.py:189-195`): keep rows that are not all zero.
    Returns (kept_index, A_sub).  A_ii = 1 always, so nothing is ever dropped."""
    keep = ~np.all(A == 0, axis=1)
    return np.where(keep)[0], A[keep][:, keep]
