#!/usr/bin/env python
"""Host-side cost of one API call (tiny inputs, metadata cached: the GPU work is negligible) next to the same
expression written with stock torch ops -- what a model pays per call when its tensors are small."""
import os
import sys
import time

import torch
from torch.nn.utils.rnn import pack_sequence, pad_sequence

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torchrua_b200 as rua  # noqa: E402

lens = torch.tensor([5, 3, 9, 1, 7, 7, 2, 4], device='cuda')
data = torch.randn((int(lens.sum()), 16), device='cuda')
c = rua.C(data=data, token_sizes=lens)
p, left = c.pack(), c.left(0)
seqs = list(torch.split(data, lens.tolist()))


def wall(label, fn, reps=2000):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    print(f'{label:34s} {(time.perf_counter() - t0) / reps * 1e6:8.1f} us per call', flush=True)


wall('C.pack()', lambda: c.pack())
wall('C.left()', lambda: c.left(0))
wall('P.cat()', lambda: p.cat())
wall('L.right()', lambda: left.right(0))
wall('C.rev()', lambda: c.rev())
wall('C.last()', lambda: c.last())
wall('C.bmask()', lambda: c.bmask())
wall('segment_sum', lambda: rua.segment_sum(data, lens))
wall('segment_logsumexp', lambda: rua.segment_logsumexp(data, lens))
wall('torch: pad_sequence', lambda: pad_sequence(seqs, batch_first=True))
wall('torch: pack_sequence', lambda: pack_sequence(seqs, enforce_sorted=False))
wall('torch: segment_reduce', lambda: torch.segment_reduce(data, 'sum', lengths=lens, unsafe=True))
wall('torch: x + 1 (one ATen op)', lambda: data + 1)
