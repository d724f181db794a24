"""PackedSequence (P) = torch.nn.utils.rnn.PackedSequence, patched in place like the reference does
(torchrua/layout/pack.py:5-55).  Token (i, t) lives at row poff[t] + unsorted_indices[i]."""
from typing import Tuple

import torch
from torch import Tensor
from torch.nn.utils.rnn import PackedSequence

from torchrua_b200 import _native

P = PackedSequence


def _ragged(self: P) -> '_native.Ragged':
    _native.require_cuda(self.data)
    return _native.ragged_from_pack(self.batch_sizes, self.sorted_indices, self.unsorted_indices,
                                    self.data.device, self.data.size()[0])


P._ragged = _ragged


def size(self: P) -> Tuple[int, ...]:
    """(B, T, *feature); batch_sizes lives on the host, so this never touches the device (pack.py:12-17)."""
    b = int(self.batch_sizes.max())
    t = self.batch_sizes.size()[0]
    return (b, t, *self.data.size()[1:])


P.size = size


def ptr(self: P) -> Tuple[Tensor, Tensor]:
    """(batch_ptr, token_ptr) in time-major order: (sorted_indices[rank], t) (pack.py:23-27)."""
    rg = self._ragged()
    token_ptr, batch_ptr, _ = _native.emit_ptr(rg.poff, self.data.size()[0], relabel=rg.sorted)
    return batch_ptr, token_ptr


P.ptr = ptr


def idx(self: P) -> P:
    n = self.data.size()[0]
    return self._replace(data=torch.arange(n, dtype=torch.long, device=self.data.device))


P.idx = idx


def offsets(self: P) -> Tensor:
    """exclusive prefix sum of batch_sizes clamped to N-1 (pack.py:43-45)."""
    rg = self._ragged()
    off, _ = _native.scan(rg.bs_dev, clamp_max=self.data.size()[0] - 1)
    return off[:-1]


P.offsets = offsets


def raw(self: P) -> Tensor:
    return self.data


P.raw = raw
