"""CPU oracle for the AutoInst NCuts hot path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and the CPU-baseline / `--impl reference` legs of
`bench.py` may import this package; the product (`autoinst_b200/`, `ncuts/`) never does.

The reference is pure Python, so the oracle is numpy/scipy:

* `affinity_ref`  – restatement of `pipeline/ncuts/ncuts_utils.py:56-67,112-113,125-156`
* `ncut_ref`      – restatement of `pipeline/ncuts/normalized_cut.py:4-63` (+ the eigsh pin of SURVEY §8c)
* `device_model`  – numpy model of the algorithm the CUDA path runs (component split + Lanczos),
                    used to check algorithm-level parity on the CPU
* `merge_ref`, `metrics_ref` – map merge and instance metrics (Appendix B of SURVEY.md)

Pinning (how this oracle is anchored): the reference has no tests or golden vectors for this path
(SURVEY.md §4, §8c).  `oracle/make_golden.py` therefore imports the UNMODIFIED reference modules
from /root/reference in the build container, checks this restatement against them on seeded inputs,
and writes the reference's own outputs to `tests/golden/*.npz`; the CPU tests replay those fixtures.
Third-party arithmetic on the path: SciPy (`cdist`, `eigsh` = ARPACK shift-invert + SuperLU), not
pinned by the reference (`setup.sh:9-10` pins numpy/open3d only); here scipy 1.18.1 / numpy 2.3.5.
"""
