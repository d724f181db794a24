#!/usr/bin/env python
"""Row-size sweep: conversions and segment reductions from 16-byte to 8 KB rows at a fixed payload (~2 GB), to
expose cliffs between the narrow-row kernels (< 128 B), the shared-warp regime and the wide rows of the bench."""
import json
import os
import statistics
import sys

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


def timed(fn, reps=9):
    ts = []
    for it in range(reps + 2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        del out
        if it >= 2:
            ts.append(a.elapsed_time(b))
    return statistics.median(ts)


ONLY = set(sys.argv[2].split(',')) if len(sys.argv) > 2 else None      # e.g. 'bwd max,seg_max'


def main():
    out = []
    print(f'{"row bytes":>9} {"B":>8} {"N":>10} | ' + ' | '.join(f'{k:>14}' for k in ('C->P', 'P->C', 'C->L', 'L->C', 'C.rev', 'seg_sum', 'seg_max', 'seg_lse', 'bwd sum', 'bwd max')))
    for hidden in (8, 16, 32, 64, 128, 256, 512, 1024, 4096):
        row = hidden * 2
        g = torch.Generator().manual_seed(0)
        target_tokens = min(int(1.5e9 // row), 48_000_000)
        b = max(target_tokens // 256, 64)
        lens = torch.randint(1, 513, (b,), generator=g)
        n, t = int(lens.sum()), int(lens.max())
        data = torch.randn((n, hidden), device='cuda', dtype=torch.bfloat16)
        c = rua.C(data=data, token_sizes=lens.cuda())
        p, left = c.pack(), c.left(0)      # metadata is cached from here on: the numbers below are kernel + dispatch
        nd, btd = n * row, b * t * row
        res = {}
        leaf = data.detach().requires_grad_(True)

        def bwd(op):
            out = getattr(rua, 'segment_' + op)(leaf, c.token_sizes)
            gout = torch.ones_like(out)
            return lambda: torch.autograd.grad(out, leaf, gout, retain_graph=True)
        for name, fn, nbytes in (('C->P', lambda: c.pack(), 2 * nd), ('P->C', lambda: p.cat(), 2 * nd),
                                 ('C->L', lambda: c.left(0), nd + btd), ('L->C', lambda: left.cat(), 2 * nd),
                                 ('C.rev', lambda: c.rev(), 2 * nd),
                                 ('seg_sum', lambda: rua.segment_sum(data, c.token_sizes), nd + b * row),
                                 ('seg_max', lambda: rua.segment_max(data, c.token_sizes), nd + b * row),
                                 ('seg_lse', lambda: rua.segment_logsumexp(data, c.token_sizes), nd + b * row),
                                 ('bwd sum', bwd('sum'), nd + b * row), ('bwd max', bwd('max'), 3 * nd + 2 * b * row)):
            if ONLY and name not in ONLY:
                continue
            ms = timed(fn)
            res[name] = {'ms': ms, 'GBs': nbytes / ms / 1e6, 'frac': nbytes / ms / 1e6 / PEAK}
        out.append({'row_bytes': row, 'B': b, 'N': n, 'ops': res})
        print(f'{row:>9} {b:>8} {n:>10} | ' + ' | '.join(f"{r['GBs']:7.0f} ({100 * r['frac']:3.0f}%)" for r in res.values()), flush=True)
        del data, c, p, left, leaf
        _native._CACHE.clear()
    if len(sys.argv) > 1 and sys.argv[1] != '-':
        json.dump(out, open(sys.argv[1], 'w'), indent=1)


if __name__ == '__main__':
    main()
