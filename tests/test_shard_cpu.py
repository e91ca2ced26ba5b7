"""CPU: catalogue sharding and the final summary gather, world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from agnfitter_b200 import shard


def test_shard_bounds_partition():
    for n in (0, 1, 7, 100, 100001):
        for world in (1, 2, 3, 8):
            cuts = [shard.shard_bounds(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            for a, b in zip(cuts, cuts[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_bounds(10, 2, 2)


def test_shard_catalog_keeps_z_slices_whole():
    rng = np.random.default_rng(0)
    z = np.round(rng.uniform(0.01, 6, size=5000), 2)          # heavy duplication of z
    seen = []
    for world in (1, 2, 4, 8):
        owned = []
        for r in range(world):
            order, lo, hi = shard.shard_catalog(z, r, world)
            owned.append(order[lo:hi])
            assert np.all(np.diff(z[order]) >= 0)
        allidx = np.concatenate(owned)
        assert np.array_equal(np.sort(allidx), np.arange(len(z)))      # a partition
        zsets = [set(z[o].tolist()) for o in owned]
        for i in range(world):
            for j in range(i + 1, world):
                assert not (zsets[i] & zsets[j]), "a z-slice is split between two ranks"
        seen.append(owned)
    # degenerate: one z for everybody -> everything lands on one rank, nothing is lost
    order, lo, hi = shard.shard_catalog(np.full(10, 0.5), 1, 2)
    order0, lo0, hi0 = shard.shard_catalog(np.full(10, 0.5), 0, 2)
    assert (hi - lo) + (hi0 - lo0) == 10


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        z = np.round(np.random.default_rng(1).uniform(0.01, 6, size=1001), 3)
        order, lo, hi = shard.shard_catalog(z, rank, world)
        mine = order[lo:hi]
        # per-source summary = [source index, z, 5 fake percentiles]
        rows = np.stack([mine.astype(np.float64), z[mine]] + [z[mine] * k for k in range(5)], axis=1)
        out = shard.gather_summaries(torch.from_numpy(rows))
        if rank == 0:
            q.put(out.numpy())
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


def test_gather_summaries_gloo_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    z = np.round(np.random.default_rng(1).uniform(0.01, 6, size=1001), 3)
    assert got.shape == (1001, 7)
    idx = got[:, 0].astype(np.int64)
    assert np.array_equal(np.sort(idx), np.arange(1001))           # every source exactly once
    assert np.array_equal(got[:, 1], z[idx])
    assert np.all(np.diff(got[:, 1]) >= 0)                         # rank blocks arrive in z order
    assert np.array_equal(got[:, 4], z[idx] * 2)


def test_gather_without_process_group_is_identity():
    t = torch.arange(6, dtype=torch.float64).reshape(3, 2)
    assert shard.gather_summaries(t) is t
