#!/usr/bin/env python
"""Does the C <-> P transpose of 8-byte rows suffer from the ACCESS PATTERN (32 neighbours in sorted order are scattered
over the whole C buffer) or from something inside the kernel?  Same kernel, same byte count, three length distributions:
random (ranks scattered), already sorted by length (rank r = sequence r: perfectly local), constant length."""
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torchrua_b200 as rua  # noqa: E402
from torchrua_b200 import _native  # noqa: E402


def timed(fn, reps=10):
    out = []
    for it in range(reps + 2):
        _native.PROFILE = []
        fn()
        torch.cuda.synchronize()
        prof, _native.PROFILE = _native.PROFILE, None
        if it >= 2:
            out.append(sum(a.elapsed_time(b) for _, a, b, _ in prof))
    return statistics.median(out)


g = torch.Generator().manual_seed(0)
base = torch.randint(1, 65, (1_000_000,), generator=g)
for name, lens in (('random U[1,64]', base), ('sorted descending', torch.sort(base, descending=True)[0]),
                   ('constant 32', torch.full_like(base, 32))):
    n = int(lens.sum())
    c = rua.C(data=torch.arange(n, device='cuda'), token_sizes=lens.cuda())
    p = c.pack()
    nbytes = 16 * n
    t_cp, t_pc = timed(lambda: c.pack()), timed(lambda: p.cat())
    print(f'{name:20s} N={n}  C->P {t_cp * 1e3:7.1f} us = {nbytes / t_cp / 1e6:6.0f} GB/s   P->C {t_pc * 1e3:7.1f} us = {nbytes / t_pc / 1e6:6.0f} GB/s', flush=True)
