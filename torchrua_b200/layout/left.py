"""LeftAlignedSequence (L): data (B, T, *), token_sizes (B,).  Token (i, t) lives at flat row i*T + t.
reference: torchrua/layout/left.py:9-87."""
from collections import namedtuple
from typing import Tuple

import torch
from torch import Tensor

from torchrua_b200 import _native
from torchrua_b200.layout._base import TokenSizesOps
from torchrua_b200.layout.cat import C


class LeftAlignedSequence(TokenSizesOps, namedtuple('LeftAlignedSequence', ['data', 'token_sizes'])):
    __slots__ = ()
    _right_aligned = False

    def size(self) -> Tuple[int, ...]:
        """(B, T, *feature) with T = max(token_sizes), as in the reference (left.py:61-66)."""
        rg = self._ragged()
        return (rg.B, rg.T, *self.data.size()[2:])

    def ptr(self) -> Tuple[Tensor, Tensor]:
        rg = self._ragged()
        batch_ptr, token_ptr, _ = _native.emit_ptr(rg.off, rg.N)
        return batch_ptr, token_ptr

    def idx(self) -> C:
        """flat storage row of every token, as a C (left.py:73-77 / right.py:74-79), one emit kernel."""
        rg = self._ragged()
        _, _, flat = _native.emit_ptr(rg.off, rg.N, want_which=False, want_within=False, flat_stride=rg.T,
                                      right_align=self._right_aligned)
        return C(data=flat, token_sizes=self.token_sizes)

    def offsets(self) -> Tensor:
        b, t, *_ = self.size()
        return torch.arange(b, dtype=torch.long, device=self.data.device) * t

    def raw(self) -> Tensor:
        return self.data.flatten(start_dim=0, end_dim=1)


L = LeftAlignedSequence
