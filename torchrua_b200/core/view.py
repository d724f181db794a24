"""Metadata conversion between layouts -- mirror of torchrua/core/view.py.

The reference recovers lengths by scattering a (B, T) int64 mask and summing it (view.py:11-25) and
sorts on the CPU (view.py:48).  Here: lengths of a P come from one binary search per sequence
(rua_lengths_from_pack), the permutation from a device radix sort (stable, descending), batch_sizes
from one binary search per time step.
"""
from numbers import Number
from typing import Union

import torch
from torch import Tensor

from torchrua_b200 import _native
from torchrua_b200.layout import C, L, P, R, Z
from torchrua_b200.utils import to_self


def token_sizes_of(self: Z) -> Tensor:
    if isinstance(self, P):
        return self._ragged().len
    return self.token_sizes


def get_mask(self: Z) -> Tensor:
    """(B, T) int64, 1 where t < len[i] (view.py:11-18) -- written directly by the mask kernel."""
    rg = self._ragged()
    return _native.mask(rg, rg.T, 0, 1, torch.long)


def cat_view(self: Union[L, P, R], **kwargs) -> C:
    return C(data=self.data, token_sizes=token_sizes_of(self))


C.cat_view = to_self
L.cat_view = cat_view
P.cat_view = cat_view
R.cat_view = cat_view


def left_view(self: Union[C, P, R], fill_value: Number, dtype: torch.dtype = None) -> L:
    return L(
        data=self.data.new_full(self.size(), fill_value=fill_value, dtype=dtype),
        token_sizes=token_sizes_of(self),
    )


C.left_view = left_view
L.left_view = to_self
P.left_view = left_view
R.left_view = left_view


def pack_view(self: Union[C, L, R], **kwargs) -> P:
    rg = self._ragged(want_pack=True)
    return P(
        data=self.data,
        batch_sizes=rg.bs_cpu,
        sorted_indices=rg.sorted,
        unsorted_indices=rg.unsorted,
    )


C.pack_view = pack_view
L.pack_view = pack_view
P.pack_view = to_self
R.pack_view = pack_view


def right_view(self: Union[C, L, P], fill_value: Number, dtype: torch.dtype = None) -> R:
    return R(
        data=self.data.new_full(self.size(), fill_value=fill_value, dtype=dtype),
        token_sizes=token_sizes_of(self),
    )


C.right_view = right_view
L.right_view = right_view
P.right_view = right_view
R.right_view = to_self
