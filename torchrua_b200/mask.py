"""Masks -- mirror of torchrua/mask.py; one write-only kernel (rua_mask) instead of
size() + new_full + ptr() + index_put_.  Left-aligned for every layout, R included (mask.py:10)."""
import torch

from torchrua_b200 import _native
from torchrua_b200.layout import C, L, P, R, T, Z


def mask(self: Z, zero, one, dtype: torch.dtype = None) -> T:
    rg = self._ragged()
    b, t, *_ = self.size()
    return _native.mask(rg, t, zero, one, self.data.dtype if dtype is None else dtype)


def bmask(self: Z) -> T:
    return self.mask(zero=False, one=True, dtype=torch.bool)


def fmask(self: Z) -> T:
    return self.mask(zero=torch.finfo(self.data.dtype).min, one=0, dtype=self.data.dtype)


for _cls in (C, L, P, R):
    _cls.mask = mask
    _cls.bmask = bmask
    _cls.fmask = fmask
