"""CattedSequence (C): data (N, *), token_sizes (B,).  Token (i, t) lives at row off[i] + t.
reference: torchrua/layout/cat.py:9-87."""
from collections import namedtuple
from typing import Tuple

import torch
from torch import Tensor

from torchrua_b200 import _native
from torchrua_b200.layout._base import TokenSizesOps


class CattedSequence(TokenSizesOps, namedtuple('CattedSequence', ['data', 'token_sizes'])):
    __slots__ = ()

    def size(self) -> Tuple[int, ...]:
        """(B, T, *feature) -- T = max length comes from the scan kernel's stats (cat.py:61-66)."""
        rg = self._ragged()
        return (rg.B, rg.T, *self.data.size()[1:])

    def ptr(self) -> Tuple[Tensor, Tensor]:
        """(batch_ptr, token_ptr) in sequence-major order (cat.py:68-71), one emit kernel."""
        rg = self._ragged()
        batch_ptr, token_ptr, _ = _native.emit_ptr(rg.off, self.data.size()[0])
        return batch_ptr, token_ptr

    def idx(self) -> 'CattedSequence':
        n = self.data.size()[0]
        return self._replace(data=torch.arange(n, dtype=torch.long, device=self.data.device))

    def offsets(self) -> Tensor:
        """exclusive prefix sum clamped to N-1 (cat.py:79-81); the clamp is fused into the scan."""
        _native.require_cuda(self.token_sizes)
        if self.token_sizes.size()[0] == 0:
            raise IndexError('index 0 is out of bounds for dimension 0 with size 0')
        off, _ = _native.scan(self.token_sizes, clamp_max=self.data.size()[0] - 1)
        return off[:-1]

    def raw(self) -> Tensor:
        return self.data


C = CattedSequence
