#!/usr/bin/env python
"""Per-op timings of the hot path at BASELINE.json's configs 2, 3 and 5 on one B200 (not the contract bench:
that is ../bench.py).  Every op is timed through the public API (metadata cache cleared before each call, so
scans / sorts / host syncs are inside the number) with CUDA events, median of `--reps` after 2 warm-ups;
`kernel` columns come from events recorded right around the dominant kernel launch.

Also times, where one exists, the stock ATen *payload* op the reference would end up launching on the same
GPU (advanced-index gather with a PRECOMPUTED index, `torch.segment_reduce`) -- a lower bound on the
reference's GPU cost that leaves out its 130-450 ATen calls of index construction and 18-48 host syncs.

    python benchmarks/ops.py [--cfg 2 3 5] [--reps 10] [--out gpurun_out/ops_r1.json]
"""
import argparse
import json
import os
import statistics
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torchrua_b200 as rua  # noqa: E402
from torchrua_b200 import _native  # noqa: E402

PEAK = 6526.2
try:
    PEAK = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
except Exception:
    pass


def timed(fn, reps, clear=True):
    """median (api_ms, kernel_ms) of fn()."""
    api, ker = [], []
    for it in range(reps + WARMUPS):
        if clear:
            _native._CACHE.clear()
        _native.PROFILE = []
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        prof, _native.PROFILE = _native.PROFILE, None
        del out
        if it >= WARMUPS:
            api.append(a.elapsed_time(b))
            ker.append(sum(s.elapsed_time(e) for _, s, e, _ in prof) if prof else float('nan'))
    return statistics.median(api), statistics.median(ker)


QUIET = False
WARMUPS = 5     # median of >= 20 timed calls after >= 5 warm-ups (--reps 20 is the default)
ONLY = None      # --only: time just the rows whose name contains one of these substrings


def compact(rows):
    """the per-op record bench.py embeds in its JSON line: algorithmic bytes, event-timed ms (public API call with the
    metadata cache cleared / kernel launches only) and the fraction of the measured HBM peak for both."""
    out = []
    for r in rows:
        e = {'op': r['op'], 'alg_bytes': int(r['alg_GB'] * 1e9), 'api_ms': round(r['api_ms'], 4),
             'api_frac': round(r['api_frac_of_measured_peak'], 4)}
        if r.get('kernel_GBs'):
            e['kernel_ms'] = round(r['kernel_ms'], 4)
            e['kernel_frac'] = round(r['kernel_frac_of_measured_peak'], 4)
        if 'aten_payload_ms' in r:
            e['aten_payload_ms'] = round(r['aten_payload_ms'], 4)
        out.append(e)
    return out


def row(results, cfg, name, nbytes, tokens, fn, reps, aten=None):
    if ONLY and not any(k in name for k in ONLY):
        return
    api_ms, ker_ms = timed(fn, reps)
    r = {'cfg': cfg, 'op': name, 'api_ms': api_ms, 'kernel_ms': ker_ms, 'alg_GB': nbytes / 1e9,
         'api_GBs': nbytes / api_ms / 1e6, 'kernel_GBs': nbytes / ker_ms / 1e6 if ker_ms == ker_ms else None,
         'Mtok_s': tokens / api_ms / 1e3}
    r['kernel_frac_of_measured_peak'] = r['kernel_GBs'] / PEAK if r['kernel_GBs'] else None
    r['api_frac_of_measured_peak'] = r['api_GBs'] / PEAK
    if aten is not None:
        aten_ms, _ = timed(aten, max(3, reps // 2), clear=False)
        r['aten_payload_ms'] = aten_ms
        r['speedup_vs_aten_payload'] = aten_ms / api_ms
    results.append(r)
    if QUIET:
        return
    k = f"{r['kernel_GBs']:7.0f}" if r['kernel_GBs'] else '      -'
    extra = f"  aten {r['aten_payload_ms']:8.3f} ms (x{r['speedup_vs_aten_payload']:.1f})" if aten is not None else ''
    print(f"cfg{cfg} {name:28s} api {api_ms:8.3f} ms {r['api_GBs']:7.0f} GB/s | kernel {k} GB/s "
          f"({(r['kernel_frac_of_measured_peak'] or 0) * 100:5.1f}% of {PEAK:.0f}) | {r['Mtok_s']:9.1f} Mtok/s{extra}",
          flush=True)


def cfg2(results, reps, quiet=False, light=False):
    """light: conversions, selects, masks and reductions only (the record bench.py embeds); the full table adds the list
    constructors, strict mode, scatter_* and the backward passes."""
    global QUIET
    QUIET = quiet
    g = torch.Generator().manual_seed(0)
    lens = torch.randint(1, 513, (4096,), generator=g)
    n, b, t, d = int(lens.sum()), 4096, int(lens.max()), 2048
    data = torch.randn((n, 1024), device='cuda').to(torch.bfloat16)
    c = rua.C(data=data, token_sizes=lens.cuda())
    srcs = {'C': c, 'L': c.left(0), 'R': c.right(0), 'P': c.pack()}
    nd, btd = n * d, b * t * d
    conv = {'C': lambda z: z.cat(), 'L': lambda z: z.left(0), 'R': lambda z: z.right(0), 'P': lambda z: z.pack()}
    # the index the reference would gather with, precomputed (payload-only ATen baseline)
    bp, tp = srcs['P'].ptr()
    key_cp = (c.offsets()[bp] + tp)
    for sk in 'CLPR':
        for dk in 'CLPR':
            if sk == dk:
                continue
            nbytes = (nd + btd + 8 * b) if dk in 'LR' else (2 * nd + 8 * b + (8 * t if 'P' in (sk, dk) else 0))
            aten = (lambda: torch.index_select(data, 0, key_cp)) if (sk, dk) == ('C', 'P') else None
            row(results, 2, f'{sk}->{dk}', nbytes, n, lambda s=srcs[sk], f=conv[dk]: f(s), reps, aten)
    for sk in 'CP':
        s = srcs[sk]
        row(results, 2, f'{sk}.rev', 2 * nd, n, lambda s=s: s.rev(), reps)
        row(results, 2, f'{sk}.roll(1)', 2 * nd, n, lambda s=s: s.roll(1), reps)
        row(results, 2, f'{sk}.last', 2 * b * d + 8 * b, n, lambda s=s: s.last(), reps)
    row(results, 2, 'C.head(1)', 2 * b * d, n, lambda: c.head(1), reps)
    row(results, 2, 'C.bmask', b * t + 8 * b, n, lambda: c.bmask(), reps)
    row(results, 2, 'C.fmask', b * t * 2 + 8 * b, n, lambda: c.fmask(), reps)
    for fn in ('sum', 'max', 'logsumexp'):
        f = getattr(rua, 'segment_' + fn)
        aten = None
        if fn in ('sum', 'max'):
            lc = lens.cuda()
            aten = lambda fn=fn, lc=lc: torch.segment_reduce(data, fn, lengths=lc, unsafe=True)
        row(results, 2, f'segment_{fn}', nd + b * d + 8 * b, n, lambda f=f: f(data, c.token_sizes), reps, aten)
    if light:
        return
    # .seg(duration, segment_mean) pooling: every sequence cut into pieces of 8 tokens.  On a P the
    # reducer gathers the packed rows itself; "unfused" is the reference's composition P -> C, reduce, C -> P.
    cuts = [[8] * (l // 8) + ([l % 8] if l % 8 else []) for l in lens.tolist()]
    dur = rua.C(data=torch.tensor([x for q in cuts for x in q], device='cuda'),
                token_sizes=torch.tensor([len(q) for q in cuts], device='cuda'))
    s_rows = int(dur.data.numel())
    seg_bytes = nd + 3 * s_rows * d + 8 * s_rows          # data read, result written, result packed (read + write)
    row(results, 2, 'C.seg(8-token pieces, mean)', nd + s_rows * d + 8 * s_rows, n, lambda: c.seg(dur, rua.segment_mean), reps)
    row(results, 2, 'P.seg(8-token pieces, mean)', seg_bytes, n, lambda: srcs['P'].seg(dur, rua.segment_mean), reps)
    left = srcs['L']
    row(results, 2, 'L.seg(8-token pieces, mean)', nd + 8 * n + s_rows * (d + 8) + s_rows * d + b * int(dur.token_sizes.max()) * d, n,
        lambda: left.seg(dur, rua.segment_mean), reps)
    row(results, 2, 'L.seg unfused (reduce over all B x T rows)', btd + b * (int(dur.token_sizes.max()) + 1) * d, n,
        lambda: left.seg(dur, lambda t_, s_: rua.segment_mean(t_, s_)), reps)
    row(results, 2, 'P.seg unfused (P->C, reduce, C->P)', seg_bytes, n,
        lambda: srcs['P'].cat().seg(dur, rua.segment_mean).pack(), reps)
    # sub-word -> word pooling: very short segments (1..4 rows of 2 KB), the common use of segment_mean / .seg
    gp = torch.Generator().manual_seed(5)
    short = torch.randint(1, 5, (n,), generator=gp)
    short = short[:int(torch.searchsorted(short.cumsum(0), n))]
    short = torch.cat([short, torch.tensor([n - int(short.sum())])]).cuda()
    s_short = int(short.numel())
    for fn in ('mean', 'max', 'logsumexp'):
        f = getattr(rua, 'segment_' + fn)
        aten = (lambda fn=fn: torch.segment_reduce(data, fn, lengths=short, unsafe=True)) if fn != 'logsumexp' else None
        row(results, 2, f'segment_{fn} (pieces U[1,4])', nd + s_short * d + 8 * s_short, n, lambda f=f: f(data, short), reps, aten)
    for piece in (8, 16, 32):
        q = torch.full((n // piece,), piece)
        q = torch.cat([q, torch.tensor([n - int(q.sum())])]).cuda() if n % piece else q.cuda()
        for fn in ('mean', 'logsumexp'):
            f = getattr(rua, 'segment_' + fn)
            row(results, 2, f'segment_{fn} (pieces of {piece})', nd + int(q.numel()) * (d + 8), n, lambda f=f, q=q: f(data, q), reps)
    # constructors from a list of 4096 tensors: host-side metadata + one multi-source kernel
    from torch.nn.utils.rnn import pack_sequence, pad_sequence
    pieces = list(torch.split(data, lens.tolist()))
    row(results, 2, 'L.new(list of 4096)', nd + btd, n, lambda: rua.L.new(pieces), max(3, reps // 3),
        aten=lambda: pad_sequence(pieces, batch_first=True))
    row(results, 2, 'P.new(list of 4096)', 2 * nd, n, lambda: rua.P.new(pieces), max(3, reps // 3),
        aten=lambda: pack_sequence(pieces, enforce_sorted=False))
    row(results, 2, 'C.new(list of 4096)', 2 * nd, n, lambda: rua.C.new(pieces), max(3, reps // 3),
        aten=lambda: torch.cat(pieces, dim=0))
    del pieces
    big = list(torch.split(data, [n // 64] * 63 + [n - 63 * (n // 64)]))     # few large tensors: the multi-source kernel
    tb = n // 64 + (n - 63 * (n // 64) - n // 64)
    row(results, 2, 'L.new(list of 64 x 33 MB)', nd + 64 * max(n // 64, tb) * d, n, lambda: rua.L.new(big), reps,
        aten=lambda: pad_sequence(big, batch_first=True))
    row(results, 2, 'P.new(list of 64 x 33 MB)', 2 * nd, n, lambda: rua.P.new(big), reps,
        aten=lambda: pack_sequence(big, enforce_sorted=False))
    del big
    # parity mode: the reference's sequential, per-step-rounded accumulation (bit-identical to torch.segment_reduce)
    from torchrua_b200.reduce import strict_reductions

    def strict_sum():
        with strict_reductions():
            return rua.segment_sum(data, c.token_sizes)
    lc2 = lens.cuda()
    row(results, 2, 'segment_sum (strict mode)', nd + b * d + 8 * b, n, strict_sum, reps,
        lambda: torch.segment_reduce(data, 'sum', lengths=lc2, unsafe=True))
    # scatter_* ("next" row 8f-1): the same rows in a random order, reduced by an unsorted index
    perm = torch.randperm(n, device='cuda')
    idx = torch.repeat_interleave(torch.arange(b, device='cuda'), lens.cuda())[perm]
    src = data[perm]
    base = torch.zeros((b, 1024), device='cuda', dtype=torch.bfloat16)
    row(results, 2, 'scatter_sum (unsorted)', nd + b * d + 16 * n, n, lambda: rua.scatter_sum(base, idx, src), reps,
        lambda: torch.index_add(base, 0, idx, src))
    row(results, 2, 'scatter_max (unsorted)', nd + b * d + 16 * n, n, lambda: rua.scatter_max(base, idx, src), reps,
        lambda: torch.index_reduce(base, 0, idx, src, 'amax', include_self=False))
    del perm, idx, src, base
    # backward passes (api time of .backward() alone; the forward graph is rebuilt outside the timed region)
    leaf = data.clone().requires_grad_(True)
    lc = lens.cuda()

    def bwd(build, nbytes, name, aten_build=None):
        def make(bf):
            out = bf()
            gout = torch.ones_like(out)
            return lambda: torch.autograd.grad(out, leaf, gout, retain_graph=True)
        row(results, 2, name, nbytes, n, make(build), reps, make(aten_build) if aten_build else None)

    cl = rua.C(data=leaf, token_sizes=lc)
    bwd(lambda: cl.pack().data, 2 * nd, 'bwd C->P')
    bwd(lambda: cl.left(0).data, 2 * nd + 8 * b, 'bwd C->L')   # reads the N live rows of the padded gradient, writes N rows
    bwd(lambda: rua.segment_sum(leaf, lc), nd + b * d, 'bwd segment_sum',
        lambda: torch.segment_reduce(leaf, 'sum', lengths=lc, unsafe=True))
    bwd(lambda: rua.segment_max(leaf, lc), 3 * nd + 2 * b * d, 'bwd segment_max',
        lambda: torch.segment_reduce(leaf, 'max', lengths=lc, unsafe=True))
    bwd(lambda: rua.segment_logsumexp(leaf, lc), 2 * nd + 2 * b * d, 'bwd segment_logsumexp')
    # the same on sub-word pieces (segments of 1..4 rows)
    bwd(lambda: rua.segment_mean(leaf, short), nd + s_short * d, 'bwd segment_mean (pieces U[1,4])',
        lambda: torch.segment_reduce(leaf, 'mean', lengths=short, unsafe=True))
    bwd(lambda: rua.segment_max(leaf, short), 2 * nd + 2 * s_short * d, 'bwd segment_max (pieces U[1,4])')   # ties counted in-kernel: data read once


def cfg3(results, reps, quiet=False):
    global QUIET
    QUIET = quiet
    rng = np.random.default_rng(0)
    sizes = np.minimum(rng.zipf(1.5, 16384), 4096).astype(np.int64)
    n, s, d = int(sizes.sum()), 16384, 8192
    if not quiet:
        print(f'cfg3: B={s}, Zipf(a=1.5) clipped to 4096, N={n}, hidden 4096 bf16, payload {n * d / 1e9:.2f} GB', flush=True)
    data = torch.randn((n, 4096), device='cuda', dtype=torch.bfloat16)
    sz = torch.from_numpy(sizes).cuda()
    nd = n * d
    for fn in ('sum', 'mean', 'max', 'min', 'logsumexp', 'prod'):
        f = getattr(rua, 'segment_' + fn)
        aten = (lambda fn=fn: torch.segment_reduce(data, fn, lengths=sz, unsafe=True)) if fn in ('sum', 'max') else None
        row(results, 3, f'segment_{fn}', nd + s * d + 8 * s, n, lambda f=f: f(data, sz), reps, aten)
    c = rua.C(data=data, token_sizes=sz)
    p = c.pack()
    for name, z in (('C', c), ('P', p)):
        row(results, 3, f'{name}.head(1)', 2 * s * d, n, lambda z=z: z.head(1), reps)
        row(results, 3, f'{name}.last', 2 * s * d + 8 * s, n, lambda z=z: z.last(), reps)
        row(results, 3, f'{name}.roll(1)', 2 * nd, n, lambda z=z: z.roll(1), reps)
        row(results, 3, f'{name}.rev', 2 * nd, n, lambda z=z: z.rev(), reps)
    row(results, 3, 'C->P', 2 * nd, n, lambda: c.pack(), reps)
    row(results, 3, 'P->C', 2 * nd, n, lambda: p.cat(), reps)


def cfg5(results, reps, quiet=False):
    global QUIET
    QUIET = quiet
    g = torch.Generator().manual_seed(0)
    lens = torch.randint(1, 65, (1_000_000,), generator=g)
    n, b, t = int(lens.sum()), 1_000_000, 64
    data = torch.arange(n, dtype=torch.long, device='cuda')
    lc = lens.cuda()
    c = rua.C(data=data, token_sizes=lc)
    left = c.left(0)
    p = c.pack()
    row(results, 5, 'C.offsets', 16 * b, n, lambda: c.offsets(), reps,
        aten=lambda: torch.cumsum(lc, 0).roll(1))
    row(results, 5, 'C.bmask', b * t + 8 * b, n, lambda: c.bmask(), reps)
    row(results, 5, 'C.mask(long)', b * t * 8 + 8 * b, n, lambda: c.mask(-1, 2, torch.long), reps,
        aten=lambda: torch.full((b, t), 2, dtype=torch.long, device='cuda'))   # write-only ceiling: a plain fill
    row(results, 5, 'C.ptr', 16 * n + 8 * b, n, lambda: c.ptr(), reps,
        aten=lambda: torch.repeat_interleave(lc, output_size=n))
    row(results, 5, 'P.ptr', 16 * n + 8 * b, n, lambda: p.ptr(), reps)
    row(results, 5, 'L.idx', 8 * n + 8 * b, n, lambda: left.idx(), reps)
    row(results, 5, 'pack_view (sort+bs)', 8 * b * 3 + 8 * t, n, lambda: c.pack_view(), reps,
        aten=lambda: torch.sort(lc, descending=True, stable=True))
    row(results, 5, 'P lengths (cat_view)', 16 * b + 8 * t, n, lambda: p.cat_view(), reps)
    row(results, 5, 'C->P (8 B rows)', 16 * n + 8 * (b + t), n, lambda: c.pack(), reps)
    row(results, 5, 'C->L (8 B rows)', 8 * n + 8 * b * t + 8 * b, n, lambda: c.left(0), reps)
    row(results, 5, 'P->C (8 B rows)', 16 * n + 8 * (b + t), n, lambda: p.cat(), reps)
    row(results, 5, 'L->C (8 B rows)', 16 * n + 8 * b, n, lambda: left.cat(), reps)
    row(results, 5, 'C.rev (8 B rows)', 16 * n, n, lambda: c.rev(), reps)
    fdata = torch.randn(n, device='cuda')
    row(results, 5, 'segment_sum (H=1 fp32)', 4 * n + 4 * b + 8 * b, n, lambda: rua.segment_sum(fdata, lc), reps,
        aten=lambda: torch.segment_reduce(fdata, 'sum', lengths=lc, unsafe=True))
    row(results, 5, 'segment_max (H=1 fp32)', 4 * n + 4 * b + 8 * b, n, lambda: rua.segment_max(fdata, lc), reps)
    # backward of the per-token-scalar reductions (api time of the backward alone; graph built outside the timed region)
    leaf = fdata.clone().requires_grad_(True)

    def bwd(fn, nbytes, name):
        out = fn(leaf, lc)
        gout = torch.ones_like(out)
        row(results, 5, name, nbytes, n, lambda: torch.autograd.grad(out, leaf, gout, retain_graph=True), reps)
    bwd(rua.segment_sum, 4 * n + 4 * b + 8 * b, 'bwd segment_sum (H=1 fp32)')
    bwd(rua.segment_max, 2 * (2 * 4 * n) + 2 * 4 * b + 8 * b, 'bwd segment_max (H=1 fp32)')
    bwd(rua.segment_logsumexp, 2 * 4 * n + 2 * 4 * b + 8 * b, 'bwd segment_logsumexp (H=1 fp32)')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--cfg', type=int, nargs='+', default=[2, 3, 5])
    ap.add_argument('--reps', type=int, default=20)
    ap.add_argument('--out', default=None)
    ap.add_argument('--only', nargs='+', default=None, help='substrings of the row names to time')
    args = ap.parse_args()
    global ONLY
    ONLY = args.only
    results = []
    for k in args.cfg:
        {2: cfg2, 3: cfg3, 5: cfg5}[k](results, args.reps)
        torch.cuda.empty_cache()
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        json.dump({'peak_GBs': PEAK, 'gpu': torch.cuda.get_device_name(0), 'rows': results}, open(args.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
