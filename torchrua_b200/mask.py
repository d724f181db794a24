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


# ------------------------------------------------------------------------------------------------------------------
# fused consumer patterns (extensions -- the reference has no such calls): what a model does right
# after the hot path is `x = seq.left(0)` + `bias = seq.fmask()` (attention) or a `cu_seqlens` vector (varlen kernels).
# ------------------------------------------------------------------------------------------------------------------
def left_mask(self: Z, fill_value=0, zero=False, one=True, dtype: torch.dtype = torch.bool):
    """``(self.left(fill_value), self.mask(zero, one, dtype))`` from ONE launch: the decode of the padded destination
    also writes the mask (rua_row_map_mask).  Bit-identical to the two separate calls."""
    from torchrua_b200._lib import LEFT
    from torchrua_b200._native import MapSpec, SideSpec
    from torchrua_b200.core.cast import side_of
    from torchrua_b200.core.view import token_sizes_of
    if isinstance(self, L):
        return self, self.mask(zero=zero, one=one, dtype=dtype)
    rg = self._ragged()
    b, t = rg.B, rg.T
    spec = MapSpec(rg=rg, src=side_of(self, rg), dst=SideSpec(LEFT, width=t, rows=b * t))
    fused = _native.row_map_mask(self.raw(), spec, fill_value, zero, one, self.data.dtype if dtype is None else dtype)
    if fused is None:   # narrow rows / gradients wanted: the two native launches
        return self.left(fill_value), self.mask(zero=zero, one=one, dtype=dtype)
    data, m = fused
    data = data.view((b, t) + tuple(data.size()[1:]))
    return L(data=data, token_sizes=token_sizes_of(self)), m.view(b, t)


def left_bmask(self: Z, fill_value=0):
    return left_mask(self, fill_value, zero=False, one=True, dtype=torch.bool)


def left_fmask(self: Z, fill_value=0):
    """padded activations + the additive attention bias (0 on tokens, finfo.min on padding; mask.py:31-32)."""
    return left_mask(self, fill_value, zero=torch.finfo(self.data.dtype).min, one=0, dtype=self.data.dtype)


def cu_seqlens(self: Z) -> T:
    """(B + 1,) int32 cumulative sequence lengths, the metadata varlen attention kernels take; a by-product of the scan
    every conversion already runs (no extra pass over the lengths)."""
    return self._ragged().off.to(torch.int32)


for _cls in (C, L, P, R):
    _cls.mask = mask
    _cls.bmask = bmask
    _cls.fmask = fmask
    _cls.left_mask = left_mask
    _cls.left_bmask = left_bmask
    _cls.left_fmask = left_fmask
    _cls.cu_seqlens = cu_seqlens
