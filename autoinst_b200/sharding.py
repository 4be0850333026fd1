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


class LabelGather:
    """The label exchange of one map (or one benchmark batch), set up once and reused every pass.

    The (chunk id, length) table of every rank is exchanged ONCE here (chunk sizes are known when the inputs are
    packed, before anything is segmented).  A pass is then one `all_gather_into_tensor` from a persistent send
    buffer into a persistent receive buffer and, on the destination rank(s) only, one asynchronous copy into a
    pinned host buffer; the per-chunk results are VIEWS into that buffer (no per-chunk copies, no `.item()` syncs).
    `send` (int32, device) is where the segment call should write its labels: `segment_packed(..., dev_chunks=...)`
    takes it through `DeviceChunks.labels = gather.send_view()`.
    """

    def __init__(self, local_ids, local_sizes, num_chunks: int, *, group=None, device=None, dst=0):
        self.group = group
        self.num_chunks = int(num_chunks)
        self.local_ids = [int(i) for i in local_ids]
        self.local_sizes = [int(n) for n in local_sizes]
        self.mine = int(sum(self.local_sizes))
        self.dist = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if self.dist else 1
        self.rank = dist.get_rank(group) if self.dist else 0
        self.dst = dst                                   # rank that needs the labels on the host; None = every rank
        if device is None:
            nccl = self.dist and dist.get_backend(group) == "nccl"
            device = torch.device("cuda", torch.cuda.current_device()) if nccl else torch.device("cpu")
        self.device = torch.device(device)
        table = (self.local_ids, self.local_sizes)
        if self.dist:
            tables = [None] * self.world
            dist.all_gather_object(tables, table, group=group)          # once per map, not per pass
        else:
            tables = [table]
        self.tables = tables
        self.pad = max(max((sum(t[1]) for t in tables), default=0), 1)
        seen = sorted(i for t in tables for i in t[0])
        if seen != list(range(self.num_chunks)):
            raise ValueError("the ranks' chunk lists do not tile range(num_chunks)")
        self.send = torch.zeros(self.pad, dtype=torch.int32, device=self.device)
        self.recv = torch.empty(self.world * self.pad, dtype=torch.int32, device=self.device) if self.dist else self.send
        self.is_dst = self.dst is None or self.rank == self.dst
        self.host = None
        if self.is_dst:
            self.host = torch.empty(self.world * self.pad, dtype=torch.int32)
            if self.device.type == "cuda":
                self.host = self.host.pin_memory()
        self._event = torch.cuda.Event() if self.device.type == "cuda" else None

    def send_view(self):
        """The first `mine` entries of the send buffer: hand this to the segment call as its label output."""
        return self.send[:self.mine]

    def load(self, local_labels):
        """Copy label arrays (numpy / torch, host or device) into the send buffer, in local_ids order."""
        o = 0
        for n, lab in zip(self.local_sizes, local_labels):
            t = lab if isinstance(lab, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(lab, dtype=np.int32))
            self.send[o:o + n].copy_(t.reshape(-1).to(torch.int32), non_blocking=True)
            o += n

    def load_flat(self, flat):
        """Same for labels that are already one concatenated int32 tensor (e.g. PackedChunks.labels, pinned)."""
        self.send[:self.mine].copy_(flat[:self.mine], non_blocking=True)

    def start(self):
        """Issue the collective and (on the destination) the device-to-host copy; does not block the host."""
        if self.dist:
            dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
        if self.is_dst:
            self.host.copy_(self.recv, non_blocking=True)
            if self._event is not None:
                self._event.record()

    def finish(self):
        """Wait for the copy; returns a list of `num_chunks` numpy int32 views (None on ranks that are not dst)."""
        if not self.is_dst:
            return None
        if self._event is not None:
            self._event.synchronize()
        arr = self.host.numpy()
        out = [None] * self.num_chunks
        for r, (ids, sizes) in enumerate(self.tables):
            o = r * self.pad
            for cid, n in zip(ids, sizes):
                out[cid] = arr[o:o + n]
                o += n
        return out

    def gather(self, local_labels=None):
        if local_labels is not None:
            self.load(local_labels)
        self.start()
        return self.finish()


def gather_labels(local_ids, local_labels, num_chunks: int, group=None, device=None):
    """One-shot form: all-gather the label arrays of the chunks each rank segmented.

    local_ids: chunk indices owned by this rank; local_labels: matching list of int32 arrays
    (numpy or torch).  Returns a list of `num_chunks` numpy int32 arrays, identical on every rank.
    """
    g = LabelGather(local_ids, [int(len(lab)) for lab in local_labels], num_chunks, group=group, device=device, dst=None)
    return [a.copy() for a in g.gather(local_labels)]
