"""Oracle: recursive normalized cut of one chunk (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates `pipeline/ncuts/normalized_cut.py` of the reference:
  cut_cost :4-5, ncut_cost :7-11, get_min_ncut :13-34, normalized_cut :37-63.
The recursion is unrolled onto an explicit stack; segment order is the reference's depth-first
order with the `mask` side first (`:57-59`).

`faithful=True` repeats the reference's arithmetic expression by expression, including the two
dense `D.todense()` materialisations per threshold (`:9-10`) that dominate its run time — this is
the variant `bench.py --impl reference` times.  `faithful=False` evaluates the same association sums
as `d[cut].sum()`; `tests/test_oracle.py` checks both give the reference's partitions.

`pinned_eigsh` is the determinism recipe of SURVEY.md §8c: the reference calls
`sparse.linalg.eigsh` without `v0` (`:49`), and SciPy ≥ 1.17 then seeds ARPACK from OS entropy.
"""
from __future__ import annotations

import contextlib

import numpy as np
from scipy import sparse
import scipy.sparse.linalg as sla


def _cut_weight(w, side):
    # normalized_cut.py:4-5 — half of what is left after removing both diagonal blocks
    return (np.sum(w) - np.sum(w[side][:, side]) - np.sum(w[~side][:, ~side])) / 2


def _ncut_value(w, D, d, side, faithful):
    c = _cut_weight(w, side)                                    # :8
    if faithful:
        a = D.todense()[side].sum()                             # :9
        b = D.todense()[~side].sum()                            # :10
    else:
        a = d[side].sum()
        b = d[~side].sum()
    return (c / a) + (c / b)                                    # :11


def best_threshold_cut(ev, D, d, w, num_cuts=10, faithful=False):
    """get_min_ncut (`normalized_cut.py:13-34`): (mask, cost) of the first strictly best cut."""
    best = np.inf
    lo, hi = ev.min(), ev.max()
    side_best = np.zeros_like(ev, dtype=bool)
    if np.allclose(lo, hi):                                     # :22-23
        return side_best, best
    for t in np.linspace(lo, hi, num_cuts, endpoint=False):     # :27
        side = ev > t
        val = _ncut_value(w, D, d, side, faithful)
        if val < best:                                          # :30 strict
            side_best, best = side, val
    return side_best, best


def fiedler_of_block(w):
    """Degrees and Fiedler vector of one recursion node (`normalized_cut.py:38,42-53`)."""
    n = w.shape[0]
    W = w + sparse.identity(n)                                  # :38
    d = np.array(W.sum(axis=0))[0]                              # :42
    d2 = np.reciprocal(np.sqrt(d))                              # :43
    D = sparse.diags(d)
    D2 = sparse.diags(d2)
    L = D2 * (D - W) * D2                                       # :47
    vals, vecs = sla.eigsh(L, 2, sigma=1e-10, which='LM')       # :49 (resolved at call time → pin works)
    ev = vecs[:, np.argsort(vals)[1]]                           # :51-53
    return d, D, ev, np.sort(vals)


def normalized_cut_ref(w, num_points_orig, labels, T=0.01, split_lim=0.01, *, faithful=False,
                       trace=None):
    """Same contract as the reference's `normalized_cut(w, num_points_orig, labels, T, split_lim)`.

    NB the reference forwards T but not split_lim to its children (`:57-58`), so below the root
    the limit is always the default 0.01; reproduced here.
    trace: optional list receiving one dict per visited node (n, eigenvalues, mcut, split).
    """
    out = []
    stack = [(w, labels, split_lim)]
    while stack:
        wb, lab, lim = stack.pop()
        n = wb.shape[0]
        frac = lab.shape[0] / (num_points_orig + 1e-8)          # :39
        if not (n > 2 and frac > lim):                          # :40
            out.append(lab)
            if trace is not None:
                trace.append(dict(n=n, leaf="size"))
            continue
        d, D, ev, vals = fiedler_of_block(wb)
        side, cost = best_threshold_cut(ev, D, d, wb, 10, faithful)       # :54
        if trace is not None:
            trace.append(dict(n=n, vals=vals, mcut=float(cost), split=bool(cost < T),
                              n_side=int(side.sum())))
        if cost < T:                                            # :56
            # depth first, mask side first: push the complement below the mask side
            stack.append((wb[~side][:, ~side], lab[~side], 0.01))
            stack.append((wb[side][:, side], lab[side], 0.01))
        else:
            out.append(lab)
    return out


def canonical_sign(v):
    """Sign rule shared by the pinned oracle and the device path: sum(v) >= 0."""
    return -v if v.sum() < 0 else v


@contextlib.contextmanager
def pinned_eigsh(v0_kind="ones", seed=0, flip=False):
    """Make `scipy.sparse.linalg.eigsh` deterministic for the duration of the block:
    fixed ARPACK start vector and a canonical sign for every returned vector.  Works on the
    reference module too because it resolves `sparse.linalg.eigsh` at call time."""
    orig = sla.eigsh

    def wrapped(A, k=6, **kw):
        n = A.shape[0]
        if "v0" not in kw:
            if v0_kind == "ones":
                kw["v0"] = np.ones(n)
            else:
                kw["v0"] = np.random.default_rng(seed + n).standard_normal(n)
        vals, vecs = orig(A, k, **kw)
        vecs = np.stack([canonical_sign(vecs[:, j]) for j in range(vecs.shape[1])], axis=1)
        return vals, (-vecs if flip else vecs)      # flip=True: the opposite sign, to probe sign sensitivity

    sla.eigsh = wrapped                     # sla IS scipy.sparse.linalg: one attribute, seen by all callers
    try:
        yield
    finally:
        sla.eigsh = orig


def labels_from_groups(groups, n):
    """Integer label per point from the list of index arrays (`ncuts_utils.py:181-183`)."""
    lab = np.full(n, -1, dtype=np.int32)
    for s, g in enumerate(groups):
        lab[np.asarray(g)] = s
    return lab


def same_partition(a, b):
    """Label-permutation-invariant equality of two labelings."""
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    if a.shape != b.shape:
        return False
    pairs = len(set(zip(a.tolist(), b.tolist())))
    return pairs == len(set(a.tolist())) == len(set(b.tolist()))


def residual_group_report(lab, ref, n_orig, split_lim=0.01):
    """How a labeling `lab` relates to the reference labeling `ref` on a chunk with tiny fragments.

    On a disconnected node the normalised Laplacian is block diagonal and its null vectors are
    localised on single connected components; the two eigenvalues `eigsh(..., sigma=1e-10)` returns
    (`normalized_cut.py:49-53`) are the two largest ROUNDING residues of those null vectors, so every
    degenerate split peels exactly one component (threshold k = 1) and the chain stops when what is
    left is at most `split_lim`·N points (`:39-40`): one residual leaf made of whole tiny components,
    chosen by rounding noise of the float64 matrix.  A labeling that gives every component its own
    segment therefore equals the reference up to that residual grouping.
    Returns dict(refines, groups=[(ref_label, parts, points)], residual_only): `refines` — every
    segment of `lab` lies inside one segment of `ref`; `groups` — the reference segments that hold
    more than one segment of `lab`; `residual_only` — refines and every such group is a leaf by the
    size rule (points <= split_lim·N)."""
    lab = np.asarray(lab).ravel()
    ref = np.asarray(ref).ravel()
    owner = {}
    refines = True
    for g, r in set(zip(lab.tolist(), ref.tolist())):
        if g in owner:
            refines = False
        owner[g] = r
    parts = {}
    for g, r in owner.items():
        parts[r] = parts.get(r, 0) + 1
    groups = [(int(r), int(c), int((ref == r).sum())) for r, c in sorted(parts.items()) if c > 1]
    lim = split_lim * (n_orig + 1e-8)
    return dict(refines=refines, groups=groups,
                residual_only=bool(refines and all(p <= lim for _, _, p in groups)))
