"""CPU suite (gloo, world_size 2): host-side logic of the head-sharded multi-GPU path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ultra_b200.shard import head_range, shard_heads, all_gather_heads, OverlappedStateGather, ShardedUltra


def test_head_range_partitions_exactly():
    for B in (1, 7, 8, 128, 256, 13):
        for G in (1, 2, 3, 4, 8):
            spans = [head_range(B, r, G) for r in range(G)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, batches, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for batch in batches:
            _check(rank, world, batch, out_q)
    finally:
        dist.destroy_process_group()


def _check(rank, world, batch, out_q):
    if True:
        g = torch.Generator().manual_seed(0)
        full = torch.randn(batch, 5, 4, generator=g)  # (B, N, d) identical on every rank
        local = shard_heads(full, rank, world).clone()
        # a head-local "layer": anything that does not mix heads commutes with the sharding
        layer = lambda h: torch.relu(h * 2 + 1)
        gathered, _ = all_gather_heads(layer(local), batch)
        ok1 = torch.equal(gathered, layer(full))
        og = OverlappedStateGather(batch)
        h = local
        for li in range(3):
            h = layer(h)
            og.after_layer(li, h)
        states = og.wait()
        ref = full
        ok2 = len(states) == 3
        for st in states:
            ref = layer(ref)
            ok2 = ok2 and torch.equal(st, ref)
        og_last = OverlappedStateGather(batch, layers=[2])
        h = local
        for li in range(3):
            h = layer(h)
            og_last.after_layer(li, h)
        last = og_last.wait()
        ok3 = len(last) == 1 and torch.equal(last[0], ref)
        # ShardedUltra: local heads -> scores -> ONE all-gather of the (B_local, N) score rows == the unsharded scores
        model = lambda rg, gr, hh, rr: (hh.float() * 3 + rr.float()).unsqueeze(1) + torch.arange(6).float()  # noqa: E731
        h_index, r_index = torch.arange(batch) % 5, (torch.arange(batch) * 7) % 3
        ok4 = torch.equal(ShardedUltra(model)(None, None, h_index, r_index), model(None, None, h_index, r_index))
        out_q.put((rank, batch, bool(ok1), bool(ok2), bool(ok3 and ok4)))


def test_all_gather_heads_gloo_world2():
    batches = [8, 7]  # equal shards and ragged shards
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batches, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(2 * len(batches))]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted((r[0], r[1]) for r in res) == [(0, 7), (0, 8), (1, 7), (1, 8)]
    assert all(r[2] and r[3] and r[4] for r in res), res
