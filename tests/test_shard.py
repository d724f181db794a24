"""Host-side multi-GPU logic (torchrua_b200/shard.py) on CPU: partitioning properties, and the two
exchanges (lengths all-gather, output gather) across 2 gloo ranks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from torchrua_b200 import shard


def free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


@pytest.mark.parametrize('world', [2, 4, 8])
def test_balanced_partition(world):
    g = torch.Generator().manual_seed(0)
    lens = torch.randint(1, 513, (4096 * world,), generator=g)
    parts = shard.balanced_partition(lens, world)
    allids = torch.cat(parts)
    assert torch.equal(torch.sort(allids)[0], torch.arange(lens.numel())), 'every sequence owned exactly once'
    counts = [p.numel() for p in parts]
    assert max(counts) - min(counts) <= 1
    assert shard.partition_imbalance(lens, parts) < 1.001
    for p in parts:
        assert torch.equal(p, torch.sort(p)[0])
    # deterministic
    again = shard.balanced_partition(lens, world)
    assert all(torch.equal(a, b) for a, b in zip(parts, again))


def test_balanced_partition_zipf():
    rng = np.random.default_rng(0)
    lens = torch.from_numpy(np.minimum(rng.zipf(1.5, 16384), 4096))
    parts = shard.balanced_partition(lens, 8)
    assert shard.partition_imbalance(lens, parts) < 1.01


def test_take_sequences():
    lens = torch.tensor([3, 1, 4, 2])
    data = torch.arange(10)
    d, l = shard.take_sequences(data, lens, torch.tensor([0, 2]))
    assert l.tolist() == [3, 4] and d.tolist() == [0, 1, 2, 4, 5, 6, 7]


def _worker(rank, world, port, results):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(3)
        lens = torch.randint(1, 20, (37,), generator=g)
        data = torch.randn((int(lens.sum()), 5), generator=g)
        parts = shard.balanced_partition(lens, world)
        ldata, llens = shard.take_sequences(data, lens, parts[rank])
        # exchange (1): lengths
        got = shard.all_gather_lengths(llens)
        assert all(torch.equal(got[r], lens[parts[r]]) for r in range(world))
        cap = max(p.numel() for p in parts)
        buf = torch.empty((world, cap + 1), dtype=torch.long)
        shard.all_gather_lengths_fixed(llens, cap, buf)
        for r in range(world):
            n = int(buf[r, 0])
            assert n == parts[r].numel() and torch.equal(buf[r, 1:1 + n], lens[parts[r]])
        # each rank "computes" on its shard (per-sequence sums; on GPUs this is rua.segment_sum)
        off = torch.cumsum(llens, 0) - llens
        local_sum = torch.stack([ldata[o:o + n].sum(0) for o, n in zip(off.tolist(), llens.tolist())])
        # exchange (2a): per-sequence rows back in global order
        sums = shard.gather_rows_by_sequence(local_sum, parts)
        goff = torch.cumsum(lens, 0) - lens
        ref = torch.stack([data[o:o + n].sum(0) for o, n in zip(goff.tolist(), lens.tolist())])
        assert torch.allclose(sums, ref)
        # exchange (2b): per-token rows back in global C order
        full = shard.gather_catted(ldata, llens, parts, lens)
        assert torch.equal(full, data)
        results[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_exchanges_gloo_world2():
    world = 2
    port = free_port()
    with mp.Manager() as m:
        results = m.dict()
        mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
        assert all(results.get(r) for r in range(world))


def test_micro_batch_cuts_cover_every_sequence_once():
    """host-side plumbing of the overlapped output gather (bench.py: fused_step)"""
    import torch
    from torchrua_b200 import shard
    g = torch.Generator().manual_seed(0)
    for b, k in ((1, 4), (3, 8), (10, 1), (4096, 4), (4096, 7), (5, 5), (100, 3)):
        lens = torch.randint(0 if b > 3 else 1, 50, (b,), generator=g)
        cuts = shard.micro_batch_cuts(lens, k)
        assert 1 <= len(cuts) <= min(k, b)
        assert cuts[0][0] == 0 and cuts[-1][1] == b and cuts[0][2] == 0 and cuts[-1][3] == int(lens.sum())
        for (a0, a1, t0, t1), nxt in zip(cuts, cuts[1:] + [None]):
            assert a1 > a0 and t1 - t0 == int(lens[a0:a1].sum())
            if nxt is not None:
                assert nxt[0] == a1 and nxt[2] == t1
        if b >= 1000:          # balanced to within one sequence
            sizes = [t1 - t0 for _, _, t0, t1 in cuts]
            assert max(sizes) - min(sizes) <= 2 * int(lens.max())
    assert shard.micro_batch_cuts(torch.zeros(0, dtype=torch.long), 4) == []
