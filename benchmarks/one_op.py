#!/usr/bin/env python
"""Run ONE op a few times (for `ncu -k regex:...` captures).  python benchmarks/one_op.py <name> [reps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torchrua_b200 as rua  # noqa: E402

name = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
g = torch.Generator().manual_seed(0)

if name.startswith('cfg5'):
    lens = torch.randint(1, 65, (1_000_000,), generator=g).cuda()
    n = int(lens.sum())
    if name == 'cfg5_flat_sum':
        x = torch.randn(n, device='cuda')
        fn = lambda: rua.segment_sum(x, lens)
    elif name == 'cfg5_flat_max':
        x = torch.randn(n, device='cuda')
        fn = lambda: rua.segment_max(x, lens)
    elif name == 'cfg5_L_to_C':
        c = rua.C(data=torch.arange(n, device='cuda'), token_sizes=lens)
        left = c.left(0)
        fn = lambda: left.cat()
    elif name == 'cfg5_C_to_L':
        c = rua.C(data=torch.arange(n, device='cuda'), token_sizes=lens)
        fn = lambda: c.left(0)
    elif name == 'cfg5_C_to_P':
        c = rua.C(data=torch.arange(n, device='cuda'), token_sizes=lens)
        fn = lambda: c.pack()
    elif name == 'cfg5_ptr':
        c = rua.C(data=torch.arange(n, device='cuda'), token_sizes=lens)
        fn = lambda: c.ptr()
    elif name == 'cfg5_bmask':
        c = rua.C(data=torch.arange(n, device='cuda'), token_sizes=lens)
        fn = lambda: c.bmask()
    elif name == 'cfg5_all':       # every index-only / narrow-row op of config 5 once (one ncu pass over all kernels)
        c = rua.C(data=torch.arange(n, device='cuda'), token_sizes=lens)
        left, pk = c.left(0), c.pack()
        x = torch.randn(n, device='cuda')

        def fn():
            return (c.bmask(), c.mask(0, 1, torch.long), c.ptr(), pk.ptr(), left.idx(), c.pack(), c.left(0), pk.cat(),
                    left.cat(), c.rev(), rua.segment_sum(x, lens), rua.segment_max(x, lens))
    else:
        raise SystemExit(f'unknown op {name}')
elif name.startswith('cfg3'):
    sizes = np.minimum(np.random.default_rng(0).zipf(1.5, 16384), 4096).astype(np.int64)
    x = torch.randn((int(sizes.sum()), 4096), device='cuda', dtype=torch.bfloat16)
    sz = torch.from_numpy(sizes).cuda()
    op = name.split('_', 1)[1]
    fn = lambda: getattr(rua, 'segment_' + op)(x, sz)
elif name.startswith('pool'):      # sub-word -> word pooling: cfg2 payload, segments of 1..4 rows; pool_<op>
    gp = torch.Generator().manual_seed(5)
    n = 1043469
    short = torch.randint(1, 5, (n,), generator=gp)
    short = short[:int(torch.searchsorted(short.cumsum(0), n))]
    short = torch.cat([short, torch.tensor([n - int(short.sum())])]).cuda()
    x = torch.randn((n, 1024), device='cuda').to(torch.bfloat16)
    op = name.split('_', 1)[1]
    fn = lambda: getattr(rua, 'segment_' + op)(x, short)
else:
    raise SystemExit(f'unknown op {name}')

for _ in range(reps):
    out = fn()
torch.cuda.synchronize()
print('ok', name)
