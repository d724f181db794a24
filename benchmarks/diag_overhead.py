#!/usr/bin/env python
"""Where does the API time of one C->P go?  Host-side wall clock of each phase + GPU events."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torchrua_b200 as rua  # noqa: E402
from torchrua_b200 import _lib, _native  # noqa: E402

g = torch.Generator().manual_seed(0)
lens = torch.randint(1, 513, (4096,), generator=g).cuda()
n = int(lens.sum())
data = torch.randn((n, 1024), device='cuda').to(torch.bfloat16)
c = rua.C(data=data, token_sizes=lens)
lib = _lib.load()


def wall(label, fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps * 1e6
    print(f'{label:40s} {dt:9.1f} us (host wall incl. sync)', flush=True)


def meta_only():
    _native._CACHE.clear()
    rg = _native.ragged_from_lengths(lens)
    rg.ensure_pack()


def scan_only():
    _native._CACHE.clear()
    _native.ragged_from_lengths(lens)


def full():
    _native._CACHE.clear()
    c.pack()


def cached():
    c.pack()


hostbuf = torch.empty(2 + 4096, dtype=torch.long, device='cuda')


def fetch_only():
    _native.fetch(hostbuf)


def empty_alloc():
    torch.empty((n, 1024), dtype=torch.bfloat16, device='cuda')


wall('scan only (1 launch, no sync)', scan_only)
wall('fetch of 4098 int64 (pinned D2H + sync)', fetch_only)
wall('torch.empty(2 GB)', empty_alloc)
wall('meta: scan + fused + fetch', meta_only)
wall('C.pack() cold (cache cleared)', full)
wall('C.pack() warm (metadata cached)', cached)
wall('C.left() warm', lambda: c.left(0))
wall('segment_sum warm', lambda: rua.segment_sum(c.data, c.token_sizes))

# GPU-side durations of the metadata kernels
for name, fn in (('scan', scan_only), ('meta', meta_only)):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    fn()
    b.record()
    torch.cuda.synchronize()
    print(f'{name}: {a.elapsed_time(b) * 1e3:.1f} us between events', flush=True)
