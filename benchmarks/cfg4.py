#!/usr/bin/env python
"""BASELINE.json configs[3]: the full layout + reduce pipeline over 64 M tokens of hidden 2048 bf16 (262 GB of payload),
batch-sharded by sequence across N B200 with length-balanced partitioning.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P benchmarks/cfg4.py

The payload does not fit next to its outputs on few GPUs, so every rank walks its shard in MICRO-BATCHES of the same
size whatever N is (2 M tokens = 8.6 GB): the per-GPU work per micro-batch is identical at 2, 4 and 8 GPUs and the
job time is (micro-batches per rank) x (time per micro-batch).  Per micro-batch:
    C -> P -> C   (bit-exact round trip)      C -> L -> C   (bit-exact round trip)      segment_sum, segment_max
and the per-sequence reductions of ALL sequences are gathered on every rank through the peer windows
(rua_scatter_rows_multi; 250 K x 4 KB = 1 GB per reduction).  Checks at full size: both round trips are the identity,
max equals an independent torch reduction on a sample of sequences, every rank ends with the same gathered table
(checksum all-reduce).  Prints one JSON line on rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torchrua_b200 as rua  # noqa: E402
from torchrua_b200 import _native, shard  # noqa: E402

TOKENS, HIDDEN, MAX_LEN, MICRO_TOKENS = 64_000_000, 2048, 512, 2_000_000
PEAK = 6526.2
try:
    PEAK = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
except Exception:
    pass


def run(rank: int, world: int, dev, micro_batches: int = 0):
    """the whole 64 M-token job on `world` ranks (STRONG scaling: the total is fixed, each rank walks 1/world of it in
    2 M-token micro-batches).  Needs an initialised NCCL process group when world > 1.  Returns the record (identical on
    every rank).  `micro_batches` > 0 times only that many micro-batches per rank (smoke runs)."""
    g = torch.Generator().manual_seed(0)
    b_total = int(TOKENS / ((1 + MAX_LEN) / 2))
    glens = torch.randint(1, MAX_LEN + 1, (b_total,), generator=g)
    parts = shard.balanced_partition(glens, world)
    mine = parts[rank]
    lens_host = glens[mine]
    total = int(glens.sum())
    # micro-batches: consecutive runs of this rank's sequences holding <= MICRO_TOKENS tokens
    csum = torch.cumsum(lens_host, 0)
    cuts, start = [], 0
    while start < lens_host.numel():
        base = int(csum[start - 1]) if start else 0
        end = int(torch.searchsorted(csum, base + MICRO_TOKENS, right=True))
        end = max(end, start + 1)
        cuts.append((start, end))
        start = end
    if micro_batches > 0:
        cuts = cuts[:micro_batches]
    row = HIDDEN * 2
    windows = shard.PeerWindows(2 * ((b_total * row + 255) // 256 * 256)) if world > 1 else None
    max_off = (b_total * row + 255) // 256 * 256
    sums_local = torch.zeros((lens_host.numel(), HIDDEN), dtype=torch.bfloat16, device=dev)
    maxs_local = torch.zeros_like(sums_local)

    ms_total, checked, alg_bytes, tokens_done = 0.0, 0, 0, 0
    gen = torch.Generator(device=dev).manual_seed(1 + rank)
    # the first micro-batch runs twice: once untimed (allocator growth, lazy kernel loading), then for the record
    for k, (a, b) in enumerate([cuts[0]] + cuts):
        ln = lens_host[a:b].to(dev)
        n = int(lens_host[a:b].sum())
        data = torch.randn((n, HIDDEN), generator=gen, device=dev, dtype=torch.float32).to(torch.bfloat16)
        _native._CACHE.clear()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        c = rua.C(data=data, token_sizes=ln)
        back_p = c.pack().cat()
        back_l = c.left(0).cat()
        s = rua.segment_sum(data, ln)
        m = rua.segment_max(data, ln)
        sums_local[a:b] = s
        maxs_local[a:b] = m
        e1.record()
        torch.cuda.synchronize()
        if k == 0:
            continue
        ms_total += e0.elapsed_time(e1)
        tokens_done += n
        # algorithmic bytes: C->P, P->C, L->C: 2 N D each; C->L: N D + B T D; each reduction: N D + B D
        padded = (b - a) * int(lens_host[a:b].max()) * row
        alg_bytes += 6 * n * row + (n * row + padded) + 2 * (n * row + (b - a) * row)
        assert torch.equal(back_p.data, data) and torch.equal(back_l.data, data), 'round trip is not the identity'
        if k == 1:   # independent check of the reductions on a sample of sequences
            off = torch.cumsum(ln, 0) - ln
            for i in range(0, ln.numel(), max(ln.numel() // 64, 1)):
                seg = data[int(off[i]):int(off[i]) + int(ln[i])]
                assert torch.equal(m[i], seg.max(0).values)
                torch.testing.assert_close(s[i].float(), seg.float().sum(0), rtol=1e-2, atol=1e-2 * float(ln[i]) ** 0.5)
                checked += 1
        del data, c, back_p, back_l, s, m
    # gather the per-sequence results of all ranks (global sequence order) through the peer windows
    if world > 1:
        windows.fence()          # untimed: the first collective initialises NCCL
        torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    if world > 1:
        windows.fence()
        gs = shard.gather_rows_fused(sums_local, parts, windows, offset_bytes=0, fence=False)
        gm = shard.gather_rows_fused(maxs_local, parts, windows, offset_bytes=max_off, fence=False)
        windows.fence()
    else:
        gs, gm = sums_local, maxs_local
    g1.record()
    torch.cuda.synchronize()
    gather_ms = g0.elapsed_time(g1)
    assert torch.equal(gs[mine.to(dev)], sums_local) and torch.equal(gm[mine.to(dev)], maxs_local)
    chk = torch.stack([gs.view(torch.int16).long().sum(), gm.view(torch.int16).long().sum()])
    t = torch.tensor([ms_total, gather_ms, float(tokens_done)], dtype=torch.double, device=dev)
    tok = torch.tensor([float(tokens_done)], dtype=torch.double, device=dev)
    if world > 1:
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), 'ranks disagree on the gathered results'
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tok, op=dist.ReduceOp.SUM)
    pipe_ms, gather_ms, done = float(t[0]), float(t[1]), float(tok[0])
    whole = micro_batches <= 0
    record = {
        'workload': 'configs[3]: C->P->C + C->L->C + segment_sum + segment_max, 64M tokens, hidden 2048 bf16, sequence-sharded, '
                    'length-balanced; per-sequence results of ALL sequences gathered on every rank',
        'scaling': 'strong', 'n_gpus': world, 'tokens_total': total, 'tokens_timed': int(done), 'whole_job': whole,
        'sequences_total': b_total, 'micro_batches_per_rank': len(cuts), 'warmup': 'first micro-batch run once untimed',
        'micro_batch_tokens': MICRO_TOKENS, 'imbalance': shard.partition_imbalance(glens, parts),
        'pipeline_ms_max_over_ranks': pipe_ms, 'gather_ms': gather_ms,
        'tokens_per_s': done / (pipe_ms * 1e-3), 'tokens_per_s_with_gather': done / ((pipe_ms + gather_ms) * 1e-3),
        'algorithmic_GBs_per_gpu': alg_bytes / (ms_total * 1e-3) / 1e9,
        'frac_of_measured_peak_per_gpu': alg_bytes / (ms_total * 1e-3) / 1e9 / PEAK,
        'timing': 'CUDA events around every micro-batch (data generation excluded), summed; max over ranks',
        'checks': {'round_trips_identity': True, 'reductions_sampled': checked, 'gathered_tables_identical': True},
    }
    if windows is not None:
        windows.close()
    del sums_local, maxs_local, gs, gm
    torch.cuda.empty_cache()
    return record


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (('RANK', 0), ('WORLD_SIZE', 1), ('LOCAL_RANK', 0)))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    record = run(rank, world, dev)
    if rank == 0:
        print(json.dumps(record))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
