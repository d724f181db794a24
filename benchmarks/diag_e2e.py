#!/usr/bin/env python
"""Timeline of the three-stream e2e pipeline (diagnostic)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torchrua_b200 as rua
from torchrua_b200 import _native

dev = torch.device('cuda', 0)
g = torch.Generator().manual_seed(0)
lens_host = torch.randint(1, 513, (4096,), generator=g)
n = int(lens_host.sum())
h_data = torch.randn((n, 1024)).to(torch.bfloat16).pin_memory()
h_lens = lens_host.pin_memory()
h_back = torch.empty_like(h_data).pin_memory()
s_in, s_comp, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
d_in = [torch.empty((n, 1024), dtype=torch.bfloat16, device=dev) for _ in range(2)]
d_len = [torch.empty(4096, dtype=torch.long, device=dev) for _ in range(2)]
ev_in = [torch.cuda.Event() for _ in range(2)]
ev_free = [torch.cuda.Event() for _ in range(2)]
for e in ev_free: e.record()
marks = []
t_host0 = time.perf_counter()

def step(d, ln):
    _native._CACHE.clear()
    c = rua.C(data=d, token_sizes=ln)
    p = c.pack(); l = p.left(0); r = l.right(0); back = r.cat()
    return back

def it(k):
    b = k & 1
    h0 = time.perf_counter() - t_host0
    with torch.cuda.stream(s_in):
        s_in.wait_event(ev_free[b])
        a0 = torch.cuda.Event(enable_timing=True); a0.record(s_in)
        d_in[b].copy_(h_data, non_blocking=True); d_len[b].copy_(h_lens, non_blocking=True)
        a1 = torch.cuda.Event(enable_timing=True); a1.record(s_in)
        ev_in[b].record(s_in)
    h1 = time.perf_counter() - t_host0
    with torch.cuda.stream(s_comp):
        s_comp.wait_event(ev_in[b])
        c0 = torch.cuda.Event(enable_timing=True); c0.record(s_comp)
        back = step(d_in[b], d_len[b])
        c1 = torch.cuda.Event(enable_timing=True); c1.record(s_comp)
        ev_free[b].record(s_comp)
    h2 = time.perf_counter() - t_host0
    with torch.cuda.stream(s_out):
        s_out.wait_event(c1)
        o0 = torch.cuda.Event(enable_timing=True); o0.record(s_out)
        h_back.copy_(back.data, non_blocking=True)
        back.data.record_stream(s_out)
        o1 = torch.cuda.Event(enable_timing=True); o1.record(s_out)
    h3 = time.perf_counter() - t_host0
    marks.append((k, a0, a1, c0, c1, o0, o1, h0, h1, h2, h3))

it(0); torch.cuda.synchronize()
marks.clear()
base = torch.cuda.Event(enable_timing=True); base.record(); torch.cuda.synchronize()
t_host0 = time.perf_counter()
for k in range(6): it(k)
torch.cuda.synchronize()
for k, a0, a1, c0, c1, o0, o1, h0, h1, h2, h3 in marks:
    f = lambda e: base.elapsed_time(e)
    print(f'step {k}: H2D {f(a0):7.1f}-{f(a1):7.1f}  comp {f(c0):7.1f}-{f(c1):7.1f}  D2H {f(o0):7.1f}-{f(o1):7.1f} | host enq in {h0*1e3:7.1f} comp {h1*1e3:7.1f}-{h2*1e3:7.1f} out {h3*1e3:7.1f}')
