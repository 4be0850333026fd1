"""CPU tests of the multi-rank path: chunk sharding and the label all-gather over gloo (world_size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from autoinst_b200 import sharding


def test_shard_chunks_balances_quadratic_cost():
    sizes = [12000, 3000, 8000, 8000, 5000, 4000, 10000, 3500]
    for world in (1, 2, 4, 8):
        parts = sharding.shard_chunks(sizes, world)
        assert sorted(sum(parts, [])) == list(range(len(sizes)))
        load = [sum(sizes[i] ** 2 for i in p) for p in parts]
        assert max(load) <= sum(load) / world + max(s * s for s in sizes)
    assert sharding.shard_chunks(sizes, 2) == sharding.shard_chunks(sizes, 2)        # deterministic


def test_gather_without_process_group():
    out = sharding.gather_labels([1, 0], [np.array([3, 3], np.int32), np.array([0, 1, 2], np.int32)], 2)
    assert out[0].tolist() == [0, 1, 2] and out[1].tolist() == [3, 3]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, sizes, q, adjacent=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = sharding.shard_chunks(sizes, world)[rank]
        labels = [np.full(sizes[i], i, dtype=np.int32) + np.arange(sizes[i], dtype=np.int32) % 3 for i in mine]
        if adjacent:
            # consecutive slices of ONE label buffer, as bench.py and DeviceChunks.labels hand them over (single-copy path)
            flat = torch.from_numpy(np.concatenate(labels)) if labels else torch.zeros(0, dtype=torch.int32)
            offs = np.cumsum([0] + [sizes[i] for i in mine])
            labels = [flat[a:b] for a, b in zip(offs[:-1], offs[1:])]
        expect = [np.full(sizes[i], i, dtype=np.int32) + np.arange(sizes[i], dtype=np.int32) % 3 for i in range(len(sizes))]
        out = sharding.gather_labels(mine, labels, len(sizes))
        ok = all(np.array_equal(out[i], expect[i]) for i in range(len(sizes)))
        # the persistent form bench.py uses: table exchanged once, labels written straight into the send buffer, two
        # passes, the host copy on rank 0 only
        g = sharding.LabelGather(mine, [sizes[i] for i in mine], len(sizes), dst=0)
        for rep in range(2):
            view = g.send_view()
            o = 0
            for i in mine:
                view[o:o + sizes[i]] = torch.from_numpy(expect[i] + rep)
                o += sizes[i]
            g.start()
            got = g.finish()
            if rank == 0:
                ok = ok and all(np.array_equal(got[i], expect[i] + rep) for i in range(len(sizes)))
            else:
                ok = ok and got is None
        q.put((rank, ok, [len(o) for o in out]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,adjacent", [(2, False), (3, False), (2, True)])
def test_gather_labels_gloo(world, adjacent):
    sizes = [7, 120, 33, 64, 5, 90, 1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, sizes, q, adjacent)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, lens in res:
        assert ok and lens == sizes          # the gathered set is identical on every rank and shard-count invariant
