#!/usr/bin/env python
"""Benchmark of the NCuts hot path (BASELINE.json metric: NCuts chunks/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): config_tarl_spatial (alpha 1.0, theta 0.5, T 0.03) on synthetic
SemanticKITTI-shaped chunks of about 8192 major points with random 96-d TARL features.  One step =
one pass of the whole path (affinity -> degrees -> Lanczos -> N-cut scan -> partition, recursively,
-> labels) over one batch of `--batch` chunks per GPU.  Multi-GPU is weak scaling: every rank gets
its own batch of chunks; after segmenting, the ranks all-gather the label arrays (NCCL), as the map
merge needs them.

The JSON line carries: value (inputs resident in HBM), e2e (host buffers through the C ABI entry
`ancuts_segment_chunks_host`: H2D of points+features and D2H of labels inside the timed region),
roofline (matvec kernel, CUDA events around every launch, algorithmic bytes of SURVEY.md §8d),
cpu_baseline (the oracle port of the reference's CPU path on the host cores), clocks.
`--impl reference` times only the CPU oracle port (the reference is pure Python; `oracle/` is its
restatement, checked against the unmodified reference by oracle/make_golden.py).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ncuts_chunks_per_sec"
UNIT = "chunks/s"
CONFIG_NAME = "tarl_spatial"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def make_batch(batch, n_target, seed0):
    from autoinst_b200.synthetic import make_chunk
    return [make_chunk(seed0 + i, n_target=n_target, features="tarl") for i in range(batch)]


# ------------------------------------------------------------------------------------------------
# CPU side: the oracle port of the reference path, one chunk per call
# ------------------------------------------------------------------------------------------------
def _cpu_one_chunk(args):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    seed, n_target, faithful = args
    import scipy.sparse as sp
    from autoinst_b200.synthetic import CONFIGS, make_chunk
    from oracle.affinity_ref import affinity_ref, drop_isolated
    from oracle.ncut_ref import normalized_cut_ref
    cfg = CONFIGS[CONFIG_NAME]
    ch = make_chunk(seed, n_target=n_target, features="tarl")
    t0 = time.perf_counter()
    A = affinity_ref(ch.points, ch.tarl, None, alpha=cfg["alpha"], theta=cfg["theta"], gamma=cfg["gamma"])
    _keep, A = drop_isolated(A)
    w = sp.csr_matrix(A)
    groups = normalized_cut_ref(w, ch.n, np.arange(ch.n), T=cfg["T"], split_lim=0.01, faithful=faithful)
    return time.perf_counter() - t0, ch.n, len(groups)


def cpu_sample(seeds, n_target, workers, faithful=True):
    """Run the oracle port over the given chunk seeds with `workers` processes; returns (wall s, points)."""
    t0 = time.perf_counter()
    if workers <= 1:
        res = [_cpu_one_chunk((s, n_target, faithful)) for s in seeds]
    else:
        import multiprocessing as mp
        with mp.get_context("spawn").Pool(workers) as pool:
            res = pool.map(_cpu_one_chunk, [(s, n_target, faithful) for s in seeds])
    wall = time.perf_counter() - t0
    return wall, sum(r[1] for r in res), [r[0] for r in res]


def host_workers(n_target):
    cores = os.cpu_count() or 1
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 64 << 30
    per_worker = 10 * 8 * (n_target * 1.15) ** 2          # ~10 live dense float64 N x N arrays
    return int(max(1, min(cores, avail * 0.6 // per_worker, 16)))


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu_index)], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self, t_begin=None, t_end=None):
        """Median SM clock over the samples taken inside [t_begin, t_end] (wall clock); all samples if none fall inside."""
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    ts = datetime.datetime.strptime(p[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    rows.append((ts, float(p[1]), float(p[2]), float(p[3]) if p[3].replace(".", "").isdigit() else 0.0, p[5:9]))
                except ValueError:
                    continue
            os.unlink(self.path)
        except Exception:
            pass
        inside = [r for r in rows if t_begin is not None and t_begin - 0.05 <= r[0] <= t_end + 0.05]
        use = inside if inside else rows
        reasons = set()
        for r in use:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if use:
            out.update(sm_mhz=float(np.median([r[1] for r in use])), sm_max_mhz=float(max(r[2] for r in use)),
                       reasons=sorted(reasons), samples=len(use), samples_in_timed_region=len(inside),
                       power_w_max=float(max(r[3] for r in use)))
        return out


TRAFFIC_FILE = "r1c_cluster_dram_b128.csv"        # written by tools/gpu_final_s3.sh (ncu, same build)


def measured_traffic(batch, n_target, launches_per_step):
    """DRAM bytes per timed launch of the dominant kernel from the committed ncu list (profiles/r1c_cluster_dram_b128.csv:
    two passes over 128 chunks of n_target 8192), or None when the workload differs."""
    path = os.path.join(ROOT, "profiles", TRAFFIC_FILE)
    if batch != 128 or n_target != 8192 or not os.path.exists(path) or launches_per_step <= 0:
        return None
    total = 0.0
    for line in open(path):
        f = line.strip().split(",")
        if len(f) >= 7 and f[0].startswith("k_lanczos_cluster"):      # the kernel name itself contains a comma
            total += (float(f[-2]) + float(f[-1])) * 1e9
    return total / 2.0 / launches_per_step if total else None


def workload_string(n_target):
    from autoinst_b200.synthetic import CONFIGS
    cfg = CONFIGS[CONFIG_NAME]
    return (f"config_{CONFIG_NAME} NCuts (alpha {cfg['alpha']}, theta {cfg['theta']}, T {cfg['T']}), "
            f"synthetic SemanticKITTI-shaped chunks n_target={n_target}, 96-d TARL features")


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation (oracle port) on the host cores."""
    if rank != 0:
        return
    workers = host_workers(args.n_target)
    per_step = workers if args.ref_chunks <= 0 else args.ref_chunks
    seeds = [args.seed + i for i in range(per_step)]
    for _ in range(args.warmup):
        cpu_sample(seeds[:1], min(args.n_target, 1024), 1)          # warm-up: imports, page-in (tiny chunk)
    walls, pts = [], 0
    for _ in range(args.steps):
        wall, p, _each = cpu_sample(seeds, args.n_target, workers)
        walls.append(wall)
        pts = p
    ms = 1e3 * float(np.mean(walls))
    value = per_step / (ms / 1e3)
    sample = (f"{per_step} chunk(s) of n_target={args.n_target} ({pts} major points) per step, oracle port with the "
              f"reference's dense ncut_cost (faithful=True), {workers} worker process(es), 1 thread each")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(args.n_target), "chunks_per_step": per_step, "points_per_sec": pts / (ms / 1e3)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_REAL_STDOUT = None


def protect_stdout():
    """Libraries (NCCL banner, torchrun) may write to fd 1; the contract is ONE JSON line on stdout.
    fd 1 is pointed at stderr for the whole run and the line is written to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=128, help="chunks per GPU per step")
    ap.add_argument("--n-target", dest="n_target", type=int, default=8192)
    ap.add_argument("--seed", type=int, default=1000)
    ap.add_argument("--cpu-chunks", type=int, default=1, help="chunks in the cpu_baseline sample (0 = skip)")
    ap.add_argument("--ref-chunks", type=int, default=0, help="--impl reference: chunks per step (0 = one per worker)")
    ap.add_argument("--no-stats", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from autoinst_b200 import api, sharding
    from autoinst_b200.synthetic import CONFIGS

    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = CONFIGS[CONFIG_NAME]
    kw = dict(alpha=cfg["alpha"], theta=cfg["theta"], gamma=cfg["gamma"], T=cfg["T"])

    chunks = make_batch(args.batch, args.n_target, args.seed + rank * args.batch)
    packed = api.PackedChunks([c.points for c in chunks], [c.tarl for c in chunks], None, theta=cfg["theta"], pin=True)
    dev_chunks = packed.to_device(dev)
    hd = api.Handle.get(dev)
    local_ids = [rank * args.batch + i for i in range(args.batch)]
    total_chunks = args.batch * world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def gather(labels_dev):
        if world > 1:
            parts = [labels_dev[a:b] for a, b in zip(packed.off[:-1], packed.off[1:])]
            sharding.gather_labels(local_ids, parts, total_chunks, device=dev)

    def step_resident():
        api.segment_packed(packed, dev_chunks=dev_chunks, **kw)
        gather(dev_chunks.labels)

    def step_e2e():
        res = api.segment_packed(packed, device=dev, **kw)
        if world > 1:
            sharding.gather_labels(local_ids, res.labels, total_chunks, device=dev)
        return res

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                 # nvidia-smi needs ~1 s before its first sample: start before the warm-up
    # warm-up (also sizes the workspace)
    for _ in range(args.warmup):
        step_resident()
    step_e2e()
    t_begin = time.time()
    # timed region 1: inputs resident in HBM.  CUDA events around every matvec launch (timing mode 2)
    # give the roofline of the dominant kernel over exactly this region.
    hd.set_stage_timing(2)
    hd.launch_count(reset=True)
    mv_bytes = mv_ms = 0.0
    mv_launches = 0
    acc_all = {s: dict(bytes=0.0, launches=0) for s in api.STAGES}

    def step_resident_acct():
        nonlocal mv_bytes, mv_ms, mv_launches
        step_resident()
        acc = hd.accounting()
        mv_bytes += acc["matvec"]["bytes"]; mv_ms += acc["matvec"]["ms"]; mv_launches += acc["matvec"]["launches"]
        for s in api.STAGES:
            acc_all[s]["bytes"] += acc[s]["bytes"]; acc_all[s]["launches"] += acc[s]["launches"]

    ms_step = timed(step_resident_acct, args.steps)
    launches = hd.launch_count(reset=True)
    hd.set_stage_timing(0)
    # timed region 2: the same steps through the host-buffer entry point
    ms_e2e = timed(step_e2e, args.steps)
    t_end = time.time()
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None

    # parity spot check + node statistics outside the timed regions
    res = api.segment_packed(packed, device=dev, want_stats=True, **kw)
    stats = res.stats
    # per-stage time shares from one fully instrumented step (events around every launch)
    hd.set_stage_timing(1)
    api.segment_packed(packed, dev_chunks=dev_chunks, **kw)
    stage_ms = {s: v["ms"] for s, v in hd.accounting().items()}
    hd.set_stage_timing(0)

    if rank == 0:
        peak, peak_src = peaks()
        value = total_chunks / (ms_step / 1e3)
        pts_total = int(packed.off[-1]) * world
        achieved = (mv_bytes / mv_launches) / ((mv_ms / mv_launches) * 1e-3) / 1e9 if mv_launches else 0.0
        cpu = None
        if args.cpu_chunks > 0:
            wall, pts, each = cpu_sample([args.seed + i for i in range(args.cpu_chunks)], args.n_target, 1)
            cpu = {"value": args.cpu_chunks / wall, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"{args.cpu_chunks} chunk(s) of the same batch (seed {args.seed}.., {pts} major points), oracle port "
                             f"with the reference's dense ncut_cost (faithful=True), single thread, {wall:.1f} s"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_string(args.n_target),
                       "chunks_per_gpu_per_step": args.batch, "points_per_step": pts_total,
                       "points_per_sec": pts_total / (ms_step / 1e3),
                       "l2": "inputs larger than L2: every step rebuilds the float32 affinity blocks of all recursion nodes "
                             "(%.1f GB per GPU, 126 MB of L2) and streams them from HBM once per Lanczos step"
                             % (acc_all["degree"]["bytes"] / max(args.steps, 1) / 1e9),
                       "arithmetic": "W float32 in HBM; vectors, degrees, dots and cut sums float64",
                       "segments_per_chunk": float(np.mean(res.num_segments)),
                       "eig_nodes_per_chunk": (len(stats) / args.batch) if stats is not None else None,
                       "lanczos_steps_per_chunk": (float(stats["steps"].sum()) / args.batch) if stats is not None else None,
                       "unconverged_nodes": int((stats["converged"] == 0).sum()) if stats is not None else None,
                       "stage_ms_one_step": stage_ms,
                       "stage_algorithmic_gb_per_step": {s: acc_all[s]["bytes"] / args.steps / 1e9 for s in api.STAGES}},
            "e2e": {"value": total_chunks / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": packed.h2d_bytes() * world, "d2h_bytes_per_step": packed.d2h_bytes() * world},
            "gpu_launches": int(launches) * world,
            "roofline": {"bound": "hbm", "kernel": "k_lanczos_cluster<C> (matvec of the persistent Lanczos kernels, one timed launch = the concurrent kernels of one recursion level)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         "traffic": measured_traffic(args.batch, args.n_target, mv_launches / max(args.steps, 1)),
                         "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum of the cluster kernels, "
                                           "profiles/" + TRAFFIC_FILE + ", per level like achieved",
                         "peak_source": peak_src,
                         "launches_timed": mv_launches, "avg_launch_us": 1e3 * mv_ms / max(mv_launches, 1),
                         "algorithmic_bytes_per_launch": mv_bytes / max(mv_launches, 1),
                         "note": "algorithmic bytes = sum over running nodes of 4 n^2 + 8 n per Lanczos step (SURVEY.md §8d), summed over "
                                 "the steps the kernels of one level take; the blocks of small nodes stay in L2, so the DRAM "
                                 "traffic is below the algorithmic bytes"},
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
