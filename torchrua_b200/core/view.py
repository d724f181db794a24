"""Metadata conversion between layouts -- mirror of torchrua/core/view.py.

The reference recovers lengths by scattering a (B, T) int64 mask and summing it (view.py:11-25) and
sorts on the CPU (view.py:48).  Here: lengths of a P come from one binary search per sequence
(rua_lengths_from_pack), the permutation from a device radix sort (stable, descending), batch_sizes
from one binary search per time step -- for batches of up to 8192 sequences all of it in ONE kernel.
"""
from numbers import Number

import torch
from torch import Tensor

from torchrua_b200 import _native
from torchrua_b200.layout import C, L, P, R, Z
from torchrua_b200.utils import to_self


def token_sizes_of(self: Z) -> Tensor:
    """lengths of any layout; a P derives them from (batch_sizes, unsorted_indices) on the device."""
    return self._ragged().len if isinstance(self, P) else self.token_sizes


def get_mask(self: Z) -> Tensor:
    """(B, T) int64, 1 where t < len[i] (view.py:11-18) -- written directly by the mask kernel."""
    rg = self._ragged()
    return _native.mask(rg, rg.T, 0, 1, torch.long)


def cat_view(self: Z, **kwargs) -> C:
    """same storage, C metadata (view.py:21-25)."""
    return C(data=self.data, token_sizes=token_sizes_of(self))


def pack_view(self: Z, **kwargs) -> P:
    """same storage, P metadata (view.py:47-58)."""
    rg = self._ragged(want_pack=True)
    return P(data=self.data, batch_sizes=rg.bs_cpu, sorted_indices=rg.sorted, unsorted_indices=rg.unsorted)


def _padded_view(kind):
    def view(self: Z, fill_value: Number, dtype: torch.dtype = None):
        # a freshly filled (B, T, *) buffer plus the lengths (view.py:34-38, 67-71)
        buffer = self.data.new_full(self.size(), fill_value=fill_value, dtype=dtype)
        return kind(data=buffer, token_sizes=token_sizes_of(self))
    return view


left_view = _padded_view(L)
left_view.__name__ = 'left_view'
right_view = _padded_view(R)
right_view.__name__ = 'right_view'

# every layout gets the four views; a view onto the layout's own kind is the identity
for _name, _own, _fn in (('cat_view', C, cat_view), ('left_view', L, left_view),
                         ('pack_view', P, pack_view), ('right_view', R, right_view)):
    for _cls in (C, L, P, R):
        setattr(_cls, _name, to_self if _cls is _own else _fn)
