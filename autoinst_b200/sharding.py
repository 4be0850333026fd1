"""Data-parallel sharding of independent chunks over ranks, and the once-per-map label gather
(SURVEY.md §8e).  The reference's chunk loop (`pipeline/run_pipeline.py:160-195`) has no
cross-iteration state, so chunks are dealt to ranks and no collective touches the data path; the
only exchange is the gather of the per-chunk int32 label arrays for the map merge and metrics
(`run_pipeline.py:197-238`).  Works with NCCL (CUDA tensors) and gloo (CPU tensors).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_chunks(sizes, world_size: int):
    """Longest-processing-time-first assignment with cost ~ N^2.  Returns one sorted index list per rank."""
    order = sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]) ** 2, i))
    load = [0] * world_size
    out = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += int(sizes[i]) ** 2
    return [sorted(x) for x in out]


def gather_labels(local_ids, local_labels, num_chunks: int, group=None, device=None):
    """All-gather the label arrays of the chunks each rank segmented.

    local_ids: chunk indices owned by this rank; local_labels: matching list of int32 arrays
    (numpy or torch).  Returns a list of `num_chunks` numpy int32 arrays, identical on every rank.
    Two collectives: sizes, then one padded all_gather_into_tensor of the concatenated labels.
    """
    if not (dist.is_available() and dist.is_initialized()):
        out = [None] * num_chunks
        for i, lab in zip(local_ids, local_labels):
            out[i] = np.asarray(lab.cpu() if isinstance(lab, torch.Tensor) else lab, dtype=np.int32)
        return out
    world = dist.get_world_size(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    # table of (chunk id, length) per rank, padded to the largest local chunk count (built on the host: one upload
    # instead of two tiny fill kernels per chunk)
    cnt = torch.tensor([len(local_ids)], dtype=torch.int64, device=device)
    cnts = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(cnts, cnt, group=group)
    max_cnt = int(cnts.max().item())
    meta_h = np.full((max(max_cnt, 1), 2), -1, dtype=np.int64)
    for j, (i, lab) in enumerate(zip(local_ids, local_labels)):
        meta_h[j, 0] = i
        meta_h[j, 1] = len(lab)
    meta = torch.from_numpy(meta_h).to(device)
    metas = torch.empty((world * meta.shape[0], 2), dtype=torch.int64, device=device)      # concatenated form
    dist.all_gather_into_tensor(metas, meta, group=group)
    metas_h = metas.cpu().numpy().reshape(world, meta.shape[0], 2)
    totals = [int(m[m[:, 0] >= 0, 1].sum()) for m in metas_h]
    pad = max(max(totals), 1)
    flat = torch.zeros(pad, dtype=torch.int32, device=device)
    mine = int(meta_h[meta_h[:, 0] >= 0, 1].sum())
    if mine:
        tens = [lab if isinstance(lab, torch.Tensor) else torch.as_tensor(np.asarray(lab, dtype=np.int32)) for lab in local_labels]
        same_dev = all(t.device == flat.device and t.dtype == torch.int32 for t in tens)
        adjacent = same_dev and all(t.is_contiguous() for t in tens) and all(
            a.data_ptr() + 4 * a.numel() == b.data_ptr() for a, b in zip(tens[:-1], tens[1:]))
        if adjacent:
            # the usual case: consecutive slices of one label buffer (DeviceChunks.labels): one copy, no per-chunk kernels
            base = tens[0]
            whole = torch.as_strided(base, (mine,), (1,)) if len(tens) > 1 else base
            flat[:mine] = whole
        else:
            flat[:mine] = torch.cat([t.reshape(-1).to(dtype=torch.int32) for t in tens]).to(device)
    allflat = torch.empty(world * pad, dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(allflat, flat, group=group)
    allflat_h = allflat.cpu().numpy().reshape(world, pad)
    out = [None] * num_chunks
    for r in range(world):
        o = 0
        for cid, ln in metas_h[r]:
            if cid < 0:
                continue
            out[int(cid)] = allflat_h[r, o:o + int(ln)].copy()
            o += int(ln)
    return out
