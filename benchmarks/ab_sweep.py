"""A/B: time a few ops of the row-size sweep with whichever torchrua_b200 is first on sys.path (cwd)."""
import statistics, sys, os
sys.path.insert(0, os.getcwd())
import torch
import torchrua_b200 as rua
from torchrua_b200 import _native
def timed(fn, reps=9):
    ts = []
    for it in range(reps + 2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); out = fn(); b.record(); torch.cuda.synchronize(); del out
        if it >= 2: ts.append(a.elapsed_time(b))
    return statistics.median(ts)
print(rua.__file__)
for hidden in (8, 16, 32, 64):
    row = hidden * 2
    g = torch.Generator().manual_seed(0)
    target_tokens = min(int(1.5e9 // row), 48_000_000)
    b = max(target_tokens // 256, 64)
    lens = torch.randint(1, 513, (b,), generator=g)
    n = int(lens.sum())
    data = torch.randn((n, hidden), device='cuda', dtype=torch.bfloat16)
    c = rua.C(data=data, token_sizes=lens.cuda())
    p = c.pack()
    nd = n * row
    res = {}
    for name, fn in (('C->P', lambda: c.pack()), ('P->C', lambda: p.cat()), ('C.rev', lambda: c.rev()), ('C.roll', lambda: c.roll(1))):
        ms = timed(fn); res[name] = 2 * nd / ms / 1e6
    print(row, ' '.join(f'{k} {v:6.0f}' for k, v in res.items()), flush=True)
    del data, c, p
    _native._CACHE.clear()
