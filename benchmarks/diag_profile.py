#!/usr/bin/env python
"""cProfile of the HOST side of small cold calls (metadata cache cleared every time): where the ~100 us of API time of
P.last / C.last / C.bmask at configs[1] go.  python benchmarks/diag_profile.py [P.last|C.last|C.bmask|P.cat]"""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torchrua_b200 as rua  # noqa: E402
from torchrua_b200 import _native  # noqa: E402

g = torch.Generator().manual_seed(0)
lens = torch.randint(1, 513, (4096,), generator=g).cuda()
data = torch.randn((int(lens.sum()), 1024), device='cuda').to(torch.bfloat16)
c = rua.C(data=data, token_sizes=lens)
p = c.pack()
ops = {'P.last': lambda: p.last(), 'C.last': lambda: c.last(), 'C.bmask': lambda: c.bmask(), 'P.cat': lambda: p.cat()}
for name in (sys.argv[1:] or list(ops)):
    fn = ops[name]

    def cold():
        _native._CACHE.clear()
        return fn()
    for _ in range(50):
        cold()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(500):
        cold()
    host = (time.perf_counter() - t0) / 500 * 1e6
    torch.cuda.synchronize()
    print(f'== {name}: host issue time {host:.1f} us per cold call', flush=True)
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(500):
        cold()
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr, stream=sys.stdout)
    st.sort_stats('tottime').print_stats(14)
