"""GPU parity, multi-process: the fused conversion + output gather over peer-memory windows (K5)
against (a) the plain single-process result and (b) the NCCL/gloo all_gather + permute path of shard.py.

Two processes.  With >= 2 GPUs each rank owns one (NCCL, stores cross NVLink); on a one-GPU box both ranks share
cuda:0 (gloo control plane, CUDA IPC mappings of the same device) so the kernels, the window plumbing and the
fence protocol are still exercised end to end."""
import datetime
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_gpus, results):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dev = torch.device('cuda', rank % n_gpus)
    torch.cuda.set_device(dev)
    if n_gpus >= world:
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev, timeout=datetime.timedelta(seconds=180))
    else:
        dist.init_process_group('gloo', rank=rank, world_size=world, timeout=datetime.timedelta(seconds=180))
    import torchrua_b200 as rua
    from torchrua_b200 import shard
    windows = None
    try:
        g = torch.Generator().manual_seed(11)
        lens = torch.randint(0, 40, (301,), generator=g)
        lens[7] = 0
        n = int(lens.sum())
        parts = shard.balanced_partition(lens, world)
        for dtype, hidden in ((torch.bfloat16, 256), (torch.float32, 33), (torch.int64, 1), (torch.float32, 0)):
            feat = (hidden,) if hidden else ()
            data = torch.randn((n,) + feat, generator=g).mul(8).to(dtype)
            ldata, llens = shard.take_sequences(data, lens, parts[rank])
            local_c = rua.C(data=ldata.to(dev), token_sizes=llens.to(dev))
            if windows is None:
                windows = shard.PeerWindows(4 << 20)
            for kind in ('cat', 'left', 'right', 'pack'):
                z = getattr(local_c, kind)()
                out, loc = shard.gather_catted_fused(z, parts, lens, windows, local_copy=True)
                assert torch.equal(out.cpu(), data), (dtype, hidden, kind)
                assert torch.equal(loc.cpu(), ldata), (dtype, hidden, kind)
                out = shard.gather_catted_fused(z, parts, lens, windows, offset_bytes=512)
                assert torch.equal(out.cpu(), data), (dtype, hidden, kind, 'offset')
            if dtype.is_floating_point and hidden:
                sums = rua.segment_sum(local_c.data, local_c.token_sizes)
                got = shard.gather_rows_fused(sums, parts, windows)
                ref = shard.gather_rows_by_sequence(sums, parts)
                assert torch.equal(got, ref), (dtype, hidden)
                whole = rua.segment_sum(data.to(dev), lens.to(dev))
                # the same sums up to the rounding of chunk-crossing segments (fp32 accumulation either way)
                torch.testing.assert_close(got.float(), whole.float(), rtol=1e-2 if dtype == torch.bfloat16 else 1e-5,
                                           atol=1e-3)
        # exchange (1) as peer stores: the (world, cap + 1) table of [count, lengths...] rows
        cap = max(p.numel() for p in parts) + 3
        for trial in range(2):                                   # second call: cached slots, one launch
            table = shard.push_lengths(lens[parts[rank]].to(dev), cap, windows, offset_bytes=1 << 20)
            windows.fence()
            host = table.cpu()
            for r in range(world):
                k = parts[r].numel()
                assert int(host[r, 0]) == k and torch.equal(host[r, 1:1 + k], lens[parts[r]]), (trial, r)
            windows.fence()
        results[rank] = True
    finally:
        if windows is not None:
            windows.close()
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_fused_gather_two_ranks():
    world = 2
    n_gpus = torch.cuda.device_count()
    port = _free_port()
    with mp.Manager() as m:
        results = m.dict()
        mp.spawn(_worker, args=(world, port, n_gpus, results), nprocs=world, join=True)
        assert all(results.get(r) for r in range(world))
