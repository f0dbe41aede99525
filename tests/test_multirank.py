"""CPU, world_size 2 over gloo: the N > 1 host logic — baseline sharding and the final gather of ranks/statistics
(the only collective of the path)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from visco_b200.shard import gather_ranks_stats, shard_baselines


def test_shards_partition_the_baselines():
    for nbl in (0, 1, 7, 28, 2080, 19701):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                off, cnt = shard_baselines(nbl, world, r)
                seen += list(range(off, off + cnt))
            assert seen == list(range(nbl))
            sizes = [shard_baselines(nbl, world, r)[1] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    assert shard_baselines(2080, 8, 3) == (780, 260)          # MeerKAT-64: 260 baselines per GPU
    with pytest.raises(ValueError):
        shard_baselines(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nbl, ncorr, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        off, cnt = shard_baselines(nbl, world, rank)
        b = torch.arange(off * ncorr, (off + cnt) * ncorr)
        ranks = (b % 13 + 1).to(torch.int32)                       # what a rank's compress call would produce
        stats = torch.stack([b.float(), b.float() * 0.5, torch.full_like(b, 10.0, dtype=torch.float32),
                             torch.ones_like(b, dtype=torch.float32)], dim=1)
        counts = [shard_baselines(nbl, world, r)[1] * ncorr for r in range(world)]
        rk, st = gather_ranks_stats(ranks, stats, counts)
        ok = (rk.shape[0] == nbl * ncorr and torch.equal(rk, (torch.arange(nbl * ncorr) % 13 + 1).to(torch.int32))
              and torch.equal(st[:, 0], torch.arange(nbl * ncorr).float()) and bool((st[:, 3] == 1).all()))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nbl", [28, 7])       # even and uneven shards
def test_gather_world_size_2_gloo(nbl):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nbl, 4, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_single_process_gather_is_identity():
    r, s = torch.ones(3, dtype=torch.int32), torch.zeros(3, 4)
    rr, ss = gather_ranks_stats(r, s)
    assert rr is r and ss is s
