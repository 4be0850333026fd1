"""GPU tests of the map-level rows N2 (merge) and N4 (metrics), through the C ABI, against the committed outputs of the
reference's own functions (tests/golden/merge.npz, metrics.npz, written by oracle/make_golden.py) and the oracles."""
import numpy as np
import pytest

from conftest import GOLDEN
from autoinst_b200.synthetic import CONFIGS, make_map
from oracle import merge_ref as M
from oracle.metrics_ref import instance_metrics as metrics_ref

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
KEYS = ("p", "r", "f1", "ap", "ap0.25", "ap0.5", "S_assoc")


def _api():
    from autoinst_b200 import api
    return api


def test_merge_matches_the_reference_golden(cuda_device):
    """Bit-exact on points and labels against merge_chunks_unite_instances2 / remove_semantics run unmodified."""
    api = _api()
    g = np.load(f"{GOLDEN}/merge.npz")
    n = int(g["n_chunks"])
    parts = [(g[f"p{i}"], g[f"l{i}"]) for i in range(n)]
    pts, lab = api.merge_chunks(parts, device=cuda_device)
    assert np.array_equal(pts, g["merged_points"])
    assert np.array_equal(lab, g["merged_labels"])
    gt = M.compact_labels(g["gt_labels"])
    cleaned = api.remove_semantics(gt, M.compact_labels(lab), device=cuda_device)
    assert np.array_equal(cleaned, g["cleaned"])


@pytest.mark.parametrize("seed,n_chunks,n_per", [(5, 4, 1500), (9, 5, 2000), (21, 6, (800, 3000))])
def test_merge_matches_oracle_on_synthetic_maps(cuda_device, seed, n_chunks, n_per):
    api = _api()
    chunks = make_map(n_chunks, n_per, seed=seed)
    rng = np.random.default_rng(seed)
    parts = []
    for c in chunks:
        # GT instances as "segments", some split in two and a few points relabelled: united, split and unmatched cases
        seg = c.instance.astype(np.int64).copy()
        split = (c.points[:, 2] > np.median(c.points[:, 2])) & (seg % 3 == 1)
        seg[split] += 1000
        noise = rng.random(c.n) < 0.02
        seg[noise] = 5000 + rng.integers(0, 3, size=int(noise.sum()))
        lab = M.globally_unique(c.chunk_id, M.canonical_labels(seg)) * (c.instance != 0)      # background stays 0
        parts.append((c.points, lab))
    ref_pts, ref_lab = M.merge_chunks_unite_instances(parts)
    pts, lab = api.merge_chunks(parts, device=cuda_device)
    assert np.array_equal(pts, ref_pts) and np.array_equal(lab, ref_lab)
    assert len(pts) < sum(c.n for c in chunks)
    gpts, glab = M.merge_unite_gt([(c.points, c.instance) for c in chunks])
    assert np.array_equal(gpts, pts)
    gt = M.compact_labels(glab)
    pred = M.compact_labels(lab)
    assert np.array_equal(api.remove_semantics(gt, pred, device=cuda_device), M.remove_semantics(gt, pred))


def test_merge_edge_cases(cuda_device):
    api = _api()
    rng = np.random.default_rng(0)
    p = rng.normal(size=(50, 3))
    # one chunk: returned as it is, duplicates included (no remove_duplicated_points call, point_cloud_utils.py:390-393)
    p1 = np.concatenate([p, p[:5]])
    l1 = np.arange(55) % 4
    pts, lab = api.merge_chunks([(p1, l1)], device=cuda_device)
    assert np.array_equal(pts, p1) and np.array_equal(lab, l1)
    # two identical chunks: the second is united with the first instance by instance and then dropped point by point
    l2 = (np.arange(50) % 3) + 1
    ref = M.merge_chunks_unite_instances([(p, l2), (p, l2 + 100)])
    got = api.merge_chunks([(p, l2), (p, l2 + 100)], device=cuda_device)
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]) and len(got[0]) == 50
    # background only
    z = np.zeros(50, dtype=np.int64)
    ref = M.merge_chunks_unite_instances([(p, z), (p + 1.0, z)])
    got = api.merge_chunks([(p, z), (p + 1.0, z)], device=cuda_device)
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])
    # far apart chunks: nothing in the crop, labels unchanged
    ref = M.merge_chunks_unite_instances([(p, l2), (p + 500.0, l2 + 100)])
    got = api.merge_chunks([(p, l2), (p + 500.0, l2 + 100)], device=cuda_device)
    assert np.array_equal(got[1], ref[1]) and len(got[0]) == 100


def test_metrics_match_the_reference_golden(cuda_device):
    """tests/golden/metrics.npz holds the outputs of the reference's own Metrics class."""
    api = _api()
    g = np.load(f"{GOLDEN}/metrics.npz")
    for case in range(4):
        for mp in (200, 20):
            got = api.instance_metrics(g[f"c{case}_all"], g[f"c{case}_pred"], g[f"c{case}_gt"], min_points=mp, device=cuda_device)
            assert np.allclose([got[k] for k in KEYS], g[f"c{case}_mp{mp}_ref"], rtol=0, atol=1e-12), (case, mp, got)


def test_metrics_match_oracle_on_random_labelings(cuda_device):
    api = _api()
    rng = np.random.default_rng(7)
    for trial in range(6):
        n = int(rng.integers(2000, 30000))
        gt = rng.integers(0, 25, size=n)
        gt = np.sort(gt) if trial % 2 else gt                     # contiguous instances or salt-and-pepper
        pred = gt.copy()
        flip = rng.random(n) < 0.25
        pred[flip] = rng.integers(0, 40, size=int(flip.sum()))
        allp = pred.copy()
        pred[rng.random(n) < 0.05] = 0
        for mp in (200, 20, 1):
            ref = metrics_ref(allp, pred, gt, min_points=mp)
            got = api.instance_metrics(allp, pred, gt, min_points=mp, device=cuda_device)
            for k in KEYS:
                assert (np.isnan(ref[k]) and np.isnan(got[k])) or abs(got[k] - ref[k]) <= 1e-12, (trial, mp, k, got[k], ref[k])


def test_map_post_pipeline_equals_oracle_pipeline(cuda_device):
    """north_star level 3 end to end on the device: labels -> global ids -> merge -> remove_semantics -> metrics
    (api.MapPost, what bench.py --workload map runs on rank 0) against the same chain through the oracles."""
    api = _api()
    cfg = CONFIGS["tarl_spatial"]
    chunks = make_map(5, 2000, seed=9)
    res = api.segment_chunks([c.points for c in chunks], [c.tarl for c in chunks], alpha=cfg["alpha"], theta=cfg["theta"],
                             T=cfg["T"], device=cuda_device)
    post = api.MapPost(chunks, device=cuda_device, min_points=20)
    got = post.merge_and_score(torch.as_tensor(np.concatenate(res.labels)))
    parts = [(c.points, M.globally_unique(c.chunk_id, M.canonical_labels(lab))) for c, lab in zip(chunks, res.labels)]
    gl = post.global_labels(torch.as_tensor(np.concatenate(res.labels))).cpu().numpy()
    assert np.array_equal(gl, np.concatenate([p[1] for p in parts]))
    pts, lab = M.merge_chunks_unite_instances(parts)
    gpts, glab = M.merge_unite_gt([(c.points, c.instance) for c in chunks])
    gt = M.compact_labels(glab)
    allp = M.compact_labels(lab)
    ref = metrics_ref(allp, M.remove_semantics(gt, allp.copy()), gt, min_points=20)
    for k in KEYS:
        assert abs(got[k] - ref[k]) <= 1e-12, (k, got[k], ref[k])
    assert ref["S_assoc"] > 0.3
