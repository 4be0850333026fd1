"""CPU checks (numpy model) of two decisions taken in the persistent eigensolver kernel, csrc/kernels_cluster.cuh:

* the step is  three-term recurrence (alpha, beta)  +  ONE classical Gram-Schmidt pass against u1 and every Lanczos vector:
  it keeps the basis as orthogonal as Gram-Schmidt twice;
* alpha needs its own cluster exchange: leaving alpha v_k to the Gram-Schmidt sweep (one barrier less per step) loses
  orthogonality within dozens of steps.
"""
import numpy as np
import pytest
from scipy.sparse import csr_matrix
from scipy.sparse.csgraph import connected_components

from autoinst_b200.synthetic import CONFIGS, make_chunk
from oracle import device_model as M
from oracle.affinity_ref import affinity_ref


def lanczos(Wd, d, variant, steps):
    n = len(d)
    s = 1.0 / np.sqrt(d)
    u1 = np.sqrt(d)
    u1 /= np.linalg.norm(u1)
    mv = lambda x: s * (Wd @ (s * x) + s * x)
    V = np.zeros((steps + 2, n))
    V[0] = u1                                           # row 0 = deflated top vector, as in the device basis
    v = M.start_vector(n)
    v -= u1 * (u1 @ v)
    V[1] = v / np.linalg.norm(v)
    alpha, beta, bprev = [], [], 0.0
    for k in range(steps):
        w = mv(V[k + 1])
        if variant == "device":                         # three-term, then one pass
            a1 = V[k + 1] @ w
            w = (w - a1 * V[k + 1]) - (bprev * V[k] if k else 0.0)
            h = V[:k + 2] @ w
            w -= V[:k + 2].T @ h
            a = a1 + h[k + 1]
        elif variant == "alpha_in_sweep":               # beta term only, alpha comes out of the sweep
            if k:
                w = w - bprev * V[k]
            h = V[:k + 2] @ w
            w -= V[:k + 2].T @ h
            a = h[k + 1]
        else:                                           # classical Gram-Schmidt twice
            ht = 0.0
            for _ in range(2):
                h = V[:k + 2] @ w
                w -= V[:k + 2].T @ h
                ht = ht + h
            a = ht[k + 1]
        b = np.linalg.norm(w)
        alpha.append(a); beta.append(b)
        V[k + 2] = w / b
        bprev = b
    G = V @ V.T
    return np.abs(G - np.eye(len(G))).max(), np.array(alpha), np.array(beta)


@pytest.fixture(scope="module")
def node():
    cfg = CONFIGS["tarl_spatial"]
    ch = make_chunk(1000, n_target=2000, features="tarl")
    A = affinity_ref(ch.points, ch.tarl, None, alpha=cfg["alpha"], theta=cfg["theta"])
    W = A.astype(np.float32).astype(np.float64)
    nc, lab = connected_components(csr_matrix(W != 0), directed=False)
    big = np.argmax(np.bincount(lab))
    idx = np.nonzero(lab == big)[0]
    Wb = W[np.ix_(idx, idx)]
    return Wb, Wb.sum(1) + 1.0


def test_three_term_plus_one_pass_is_as_orthogonal_as_two_passes(node):
    Wb, d = node
    o_dev, a_dev, b_dev = lanczos(Wb, d, "device", 60)
    o_two, a_two, b_two = lanczos(Wb, d, "cgs2", 60)
    assert o_dev < 1e-13 and o_two < 1e-13
    assert np.allclose(a_dev, a_two, rtol=0, atol=1e-12) and np.allclose(b_dev, b_two, rtol=0, atol=1e-12)


def test_alpha_needs_its_own_exchange(node):
    Wb, d = node
    o_bad, _, _ = lanczos(Wb, d, "alpha_in_sweep", 60)
    assert o_bad > 1e-8                                 # measured: 1e-6 .. 1e-2 after 60 steps
