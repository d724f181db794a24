"""The 12 directed layout conversions -- mirror of torchrua/core/cast.py, every one a single launch of
the ragged row-map kernel (rua_row_map) with the padding fill fused in.

reference: to_cat cast.py:8-16, cat_pack_to_left :19-23, right_to_left :26-32, to_pack :41-49,
cat_pack_to_right :52-56, left_to_right :59-65.
"""
from numbers import Number
from typing import Optional

from torch import Tensor

from torchrua_b200 import _native
from torchrua_b200._lib import CAT, LEFT, PACK, RIGHT
from torchrua_b200._native import MapSpec, SideSpec
from torchrua_b200.layout import C, L, P, R, Z
from torchrua_b200.utils import to_self


def side_of(self: Z, rg: '_native.Ragged') -> SideSpec:
    """describe the storage of ``self`` to the kernel."""
    if isinstance(self, C):
        return SideSpec(CAT, rows=self.data.size()[0])
    if isinstance(self, P):
        return SideSpec(PACK, rows=self.data.size()[0])
    b, w = self.data.size()[:2]
    return SideSpec(RIGHT if isinstance(self, R) else LEFT, width=w, rows=b * w)


def _check_source(src: SideSpec, rg: '_native.Ragged') -> None:
    """host-only consistency check of the metadata against the storage, whenever N / T are already known on the host
    (they are after the first conversion of a batch): lengths that describe more tokens than the storage holds would
    make the kernels read past the payload (the reference's advanced indexing raises an IndexError there)."""
    if src.layout in (CAT, PACK):
        if rg._N is not None and rg._N > src.rows:
            raise IndexError(f'torchrua_b200: token_sizes describe {rg._N} tokens but data holds {src.rows} rows')
    elif rg._T is not None and rg._T > src.width:
        raise IndexError(f'torchrua_b200: the longest sequence has {rg._T} tokens but data.size(1) is {src.width}')


def _convert(self: Z, rg: '_native.Ragged', dst: SideSpec, fill_value: Number = 0) -> Tensor:
    src = side_of(self, rg)
    _check_source(src, rg)
    spec = MapSpec(rg=rg, src=src, dst=dst)
    return _native.row_map(self.raw(), spec, fill_value)


def to_cat(self: Z) -> C:
    rg = self._ragged()
    data = _convert(self, rg, SideSpec(CAT, rows=rg.N))
    return C(data=data, token_sizes=self.token_sizes if not isinstance(self, P) else rg.len)


C.cat = to_self
L.cat = to_cat
R.cat = to_cat
P.cat = to_cat


def _to_padded(self: Z, layout: int, fill_value: Number):
    rg = self._ragged()
    b, t = rg.B, rg.T
    data = _convert(self, rg, SideSpec(layout, width=t, rows=b * t), fill_value)
    data = data.view((b, t) + tuple(data.size()[1:]))
    return data, (self.token_sizes if not isinstance(self, P) else rg.len)


def to_left(self: Z, fill_value: Number = 0) -> L:
    data, token_sizes = _to_padded(self, LEFT, fill_value)
    return L(data=data, token_sizes=token_sizes)


cat_pack_to_left = to_left
right_to_left = to_left

C.left = to_left
L.left = to_self
P.left = to_left
R.left = to_left


def to_pack(self: Z, sorted_indices: Optional[Tensor] = None) -> P:
    """``sorted_indices`` (extension): pack with an externally supplied permutation instead of the
    device sort -- parity mode against the reference's non-stable CPU sort (SURVEY.md 8c hazard 1)."""
    data = None
    if sorted_indices is None:
        early_out, early, n_rows = [], None, -1
        if isinstance(self, C):
            # The row count of the result is the data's, so the conversion does not need N, T or batch_sizes on the
            # host: it is enqueued right behind the fused metadata kernel, and the device->host round trip that
            # PackedSequence's host-side batch_sizes requires overlaps it instead of idling the GPU.
            n_rows = self.data.size()[0]

            def early(rg):
                early_out.append(_convert(self, rg, SideSpec(PACK, rows=n_rows)))
        rg = self._ragged(want_pack=True, early=early)   # small batches: one fused metadata kernel, one D2H
        if early_out and rg._spec_ok and rg.N == n_rows:
            data = early_out[0]
    else:
        rg = _native.with_injected_pack(self._ragged(), sorted_indices)
    if data is None:
        data = _convert(self, rg, SideSpec(PACK, rows=rg.N))
    return P(data=data, batch_sizes=rg.bs_cpu, sorted_indices=rg.sorted, unsorted_indices=rg.unsorted)


C.pack = to_pack
L.pack = to_pack
P.pack = to_self
R.pack = to_pack


def to_right(self: Z, fill_value: Number = 0) -> R:
    data, token_sizes = _to_padded(self, RIGHT, fill_value)
    return R(data=data, token_sizes=token_sizes)


cat_pack_to_right = to_right
left_to_right = to_right

C.right = to_right
L.right = to_right
P.right = to_right
R.right = to_self
