"""numpy model of the algorithm the CUDA path runs (TEST INFRASTRUCTURE, see oracle/__init__.py).

It is NOT the reference's algorithm restated (that is `ncut_ref`); it mirrors, step for step, what
`autoinst_b200/csrc` does on the device, so that the algorithmic choices that differ from the
reference can be checked against the reference on the CPU:

* a node that is disconnected is split into its connected components directly instead of through
  null-space eigenvectors (`normalized_cut.py:49-58` with lambda_2 = 0); DESIGN.md §"degenerate nodes"
* the Fiedler vector comes from Lanczos with full re-orthogonalisation on M = D^-1/2 (w+I) D^-1/2
  with the known top vector D^1/2·1 deflated, instead of ARPACK shift-invert (`normalized_cut.py:49`)
* all ten thresholds are costed from one pass over the block (bucket / difference array) instead
  of ten `ncut_cost` calls (`normalized_cut.py:27-32`)
* W is held in float32; every sum is float64.
"""
from __future__ import annotations

import numpy as np
from scipy.sparse import csr_matrix
from scipy.sparse.csgraph import connected_components


def start_vector(n: int) -> np.ndarray:
    """Deterministic pseudo-random start vector (same integer hash as csrc/lanczos.cuh)."""
    x = (np.arange(n, dtype=np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
    x ^= x >> np.uint64(32)
    x = (x * np.uint64(0xD6E8FEB86659FD93)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    x ^= x >> np.uint64(32)
    return (x >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) - 0.5


def tridiag_top2(alpha, beta):
    """Two largest eigenvalues of the Lanczos tridiagonal and the eigenvector of the largest."""
    k = len(alpha)
    T = np.diag(alpha) + np.diag(beta[:k - 1], 1) + np.diag(beta[:k - 1], -1)
    vals, vecs = np.linalg.eigh(T)
    th1 = vals[-1]
    th2 = vals[-2] if k > 1 else -np.inf
    return th1, th2, vecs[:, -1]


# placement of the convergence checks in the cluster kernel (csrc/kernels_cluster.cuh, CL_CHECK_*)
CHECK_FIRST, CHECK_LO, CHECK_HI, CHECK_SAFETY = 28, 2, 12, 0.7


START_DIR = np.array([1.0, 0.7, 0.4])      # csrc/kernels_lanczos.cuh::StartVec


def smooth_start(points: np.ndarray, d: np.ndarray) -> np.ndarray:
    """Start vector of a node when coordinates are at hand: D^1/2 (q - q_0) with q the projection on START_DIR,
    plus 1 % of the hash (csrc StartVec).  The device indexes the hash by position, the model by node order."""
    q = (points - points[0]) @ START_DIR
    return np.sqrt(d) * q + 0.01 * start_vector(points.shape[0])


def lanczos_fiedler(Wb32: np.ndarray, d: np.ndarray, *, tol=1e-10, check_every=0, kmax=1024,
                    stats=None, v0=None):
    """Fiedler vector of L = I - S (w+I) S, S = D^-1/2, for one connected block.

    Wb32: (n,n) float32 block of w (unit diagonal); d: float64 degrees of W = w + I.
    check_every = 0: adaptive placement of the checks as in the cluster kernel (first at step 28, then at
    0.7 x the number of steps predicted from the geometric rate between the last two residual estimates,
    clamped to [2, 12]); > 0: fixed period (the grid-wide path uses 16).
    Returns unit-norm ev with sum(ev) >= 0 and lambda_2.
    """
    n = Wb32.shape[0]
    s = 1.0 / np.sqrt(d)
    u1 = np.sqrt(d)
    u1 /= np.linalg.norm(u1)
    Wd = Wb32.astype(np.float64)

    def matvec(x):
        z = s * x
        return s * (Wd @ z + z)                    # (w + I) z

    kcap = min(kmax, n - 1)
    V = np.zeros((kcap + 1, n))
    alpha = np.zeros(kcap)
    beta = np.zeros(kcap)
    v = start_vector(n) if v0 is None else np.array(v0, dtype=np.float64)
    v -= u1 * (u1 @ v)
    v /= np.linalg.norm(v)
    V[0] = v
    k = 0
    converged = False
    th1 = th2 = 0.0
    y = None
    next_check = CHECK_FIRST if check_every <= 0 else check_every
    prev_k, prev_res = 0, 1.0
    while k < kcap:
        w = matvec(V[k])
        # full re-orthogonalisation, classical Gram-Schmidt twice, against u1 and V[0..k]
        h_total = np.zeros(k + 1)
        for _ in range(2):
            w -= u1 * (u1 @ w)
            h = V[:k + 1] @ w
            w -= V[:k + 1].T @ h
            h_total += h
        alpha[k] = h_total[k]
        b = np.linalg.norm(w)
        beta[k] = b
        k += 1
        breakdown = b < 1e-14
        if not breakdown:
            V[k] = w / b
        if breakdown or k == kcap or k == next_check:
            th1, th2, y = tridiag_top2(alpha[:k], beta[:k])
            res = abs(b * y[-1])
            gap = max(th1 - th2, 1e-300)
            if breakdown or k == kcap or res <= tol * gap:
                converged = breakdown or (k == n - 1) or res <= tol * gap
                break
            adv = check_every
            if check_every <= 0:
                adv = CHECK_HI
                if 0.0 < res < prev_res:
                    rate = np.log(res / prev_res) / (k - prev_k)
                    need = CHECK_SAFETY * np.log(tol * gap / res) / rate
                    adv = int(min(CHECK_HI, max(CHECK_LO, np.ceil(need))))
                prev_k, prev_res = k, res
            next_check = k + adv
    ev = V[:k].T @ y
    ev /= np.linalg.norm(ev)
    if ev.sum() < 0:
        ev = -ev
    if stats is not None:
        stats.append(dict(n=n, k=k, lam2=1.0 - th1, lam3=1.0 - th2, converged=converged))
    return ev, 1.0 - th1


def scan_cuts(Wb32, d, ev, num_cuts=10):
    """All `num_cuts` threshold cuts from one pass (model of csrc ncut_scan).
    Returns (best_k, best_cost, bucket) with bucket[i] = #thresholds below ev[i];
    best_k = -1 when the node cannot be cut (`normalized_cut.py:22-23`)."""
    mn, mx = ev.min(), ev.max()
    n = ev.shape[0]
    if np.allclose(mn, mx):
        return -1, np.inf, np.zeros(n, dtype=np.int32)
    step = (mx - mn) / num_cuts
    t = np.arange(num_cuts) * step + mn                 # numpy linspace(endpoint=False) arithmetic
    bucket = (ev[:, None] > t[None, :]).sum(axis=1).astype(np.int32)
    # difference array over unordered pairs: edge (i,j), b_j < b_i, is cut for k in [b_j, b_i-1]
    diff = np.zeros(num_cuts + 2)
    iu, ju = np.nonzero(np.triu(Wb32, 1))
    wv = Wb32[iu, ju].astype(np.float64)
    lo = np.minimum(bucket[iu], bucket[ju])
    hi = np.maximum(bucket[iu], bucket[ju])
    np.add.at(diff, lo, wv)
    np.add.at(diff, hi, -wv)
    cut = np.cumsum(diff)[:num_cuts]
    vol_by_bucket = np.bincount(bucket, weights=d, minlength=num_cuts + 1)
    total = vol_by_bucket.sum()
    assoc_b = np.cumsum(vol_by_bucket)[:num_cuts]       # sum d_i over bucket <= k  (mask false)
    assoc_a = total - assoc_b
    best_k, best = -1, np.inf
    for k in range(num_cuts):
        c = cut[k] / assoc_a[k] + cut[k] / assoc_b[k]
        if c < best:
            best, best_k = c, k
    return best_k, best, bucket


def components_of(Wb32):
    g = csr_matrix(Wb32 != 0)
    return connected_components(g, directed=False)


def segment_model(W32: np.ndarray, T: float, split_lim: float = 0.01, *, tol=1e-10, kmax=1024,
                  stats=None, points=None):
    """Labels (int32, one segment id per point) the device algorithm assigns for dense float32 W.
    points: (N,3) coordinates -> smooth Lanczos start vectors as in the segment calls of the library."""
    N = W32.shape[0]
    labels = np.full(N, -1, dtype=np.int32)
    next_label = 0

    def passes(n, lim):
        return n > 2 and n / (N + 1e-8) > lim

    def split_components(idx):
        ncomp, lab = components_of(W32[np.ix_(idx, idx)])
        return [idx[lab == c] for c in range(ncomp)]

    frontier = []
    root = np.arange(N)
    if passes(N, split_lim) and T > 0:
        frontier = split_components(root)
    else:
        labels[:] = 0
        return labels
    while frontier:
        nxt = []
        for idx in frontier:
            n = idx.shape[0]
            if not passes(n, 0.01):
                labels[idx] = next_label
                next_label += 1
                continue
            Wb = W32[np.ix_(idx, idx)]
            d = 1.0 + Wb.astype(np.float64).sum(axis=1)
            ev, lam2 = lanczos_fiedler(Wb, d, tol=tol, kmax=kmax, stats=stats,
                                       v0=None if points is None else smooth_start(points[idx], d))
            k, cost, bucket = scan_cuts(Wb, d, ev)
            if stats is not None:
                stats[-1].update(mcut=float(cost), best_k=int(k))
            if k >= 0 and cost < T:
                side = bucket > k
                for part in (idx[side], idx[~side]):
                    if passes(part.shape[0], 0.01):
                        nxt.extend(split_components(part))
                    else:
                        nxt.append(part)
            else:
                labels[idx] = next_label
                next_label += 1
        frontier = nxt
    return labels
