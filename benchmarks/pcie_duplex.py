#!/usr/bin/env python
"""Is host<->device traffic full duplex on this box?  (context for bench.py's e2e number)"""
import torch
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n, dtype=torch.uint8, device='cuda')
d_out = torch.empty(n, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=5):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s1.wait_event(a)
    s2.wait_event(a)
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    gb = reps * n * (int(h2d) + int(d2h)) / 1e9
    return gb / (ms * 1e-3)


run(True, True, 1)
print(f'H2D only   {run(True, False):6.1f} GB/s')
print(f'D2H only   {run(False, True):6.1f} GB/s')
print(f'both       {run(True, True):6.1f} GB/s (sum of the two directions)')
