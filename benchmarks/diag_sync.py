import sys, torch
sys.path.insert(0, '/root/repo')
import torchrua_b200 as rua
from torchrua_b200 import _native
g = torch.Generator().manual_seed(0)
lens = torch.randint(1, 513, (4096,), generator=g)
n = int(lens.sum())
data = torch.randn((n, 1024), device='cuda').to(torch.bfloat16)
ln = lens.cuda()
def step(clear):
    if clear: _native._CACHE.clear()
    c = rua.C(data=data, token_sizes=ln)
    back = c.pack().left(0).right(0).cat()
    s = rua.segment_sum(back.data, back.token_sizes); m = rua.segment_max(back.data, back.token_sizes)
    return back, s, m
for clear in (True, False, True, False):
    for _ in range(5): step(clear)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): step(clear)
    b.record(); torch.cuda.synchronize()
    print('clear' if clear else 'cached', a.elapsed_time(b) / 20)
