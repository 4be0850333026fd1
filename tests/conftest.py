import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

CONFIG_NAMES = ["spatial", "tarl_spatial", "tarl_spatial_dino"]
GOLDEN_SEEDS = [11, 12, 13]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def load_golden(seed, name):
    """(inputs, outputs) of one golden chunk: reference affinity (dense float64) and pinned labels."""
    import scipy.sparse as sp
    inp = np.load(os.path.join(GOLDEN, f"chunk_s{seed}_inputs.npz"))
    out = np.load(os.path.join(GOLDEN, f"chunk_s{seed}_{name}.npz"))
    n = inp["points"].shape[0]
    A = sp.csr_matrix((out["A_data"], out["A_indices"], out["A_indptr"]), shape=(n, n))
    return inp, out, A


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)
