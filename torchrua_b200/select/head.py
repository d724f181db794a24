"""head(n): first n tokens of every sequence (requires n <= min length) -- mirror of
torchrua/select/head.py.  C and R are one row-map launch (the reference builds 2B / 3B views on the
host and concatenates them); P and L stay pure views exactly like the reference (head.py:22-42)."""
import torch

from torchrua_b200 import _native
from torchrua_b200._lib import CAT, LEN_CONST, RIGHT
from torchrua_b200._native import MapSpec, SideSpec
from torchrua_b200.core.cast import side_of
from torchrua_b200.layout import C, L, P, R


def cat_head(self: C, n: int) -> C:
    rg = self._ragged()
    spec = MapSpec(rg=rg, src=side_of(self, rg), dst=SideSpec(CAT, xform=LEN_CONST, arg=n, rows=rg.B * n))
    return C(data=_native.row_map(self.data, spec), token_sizes=torch.full_like(self.token_sizes, fill_value=n))


C.head = cat_head


def pack_head(self: P, n: int) -> P:
    data, batch_sizes, sorted_indices, unsorted_indices = self
    return P(
        data=data[:int(batch_sizes[0]) * n],
        batch_sizes=batch_sizes[:n],
        sorted_indices=sorted_indices,
        unsorted_indices=unsorted_indices,
    )


P.head = pack_head


def left_head(self: L, n: int) -> L:
    return L(data=self.data[:, :n], token_sizes=torch.full_like(self.token_sizes, n))


L.head = left_head


def right_head(self: R, n: int) -> R:
    rg = self._ragged()
    b = rg.B
    spec = MapSpec(rg=rg, src=side_of(self, rg), dst=SideSpec(RIGHT, xform=LEN_CONST, arg=n, width=n, rows=b * n))
    data = _native.row_map(self.raw(), spec)
    return R(data=data.view((b, n) + tuple(self.data.size()[2:])), token_sizes=torch.full_like(self.token_sizes, n))


R.head = right_head
