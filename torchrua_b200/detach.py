"""Back to Python lists -- mirror of torchrua/detach.py (host-side views; outside the hot path)."""
from typing import List, Union

import torch
from torch.types import Number

from torchrua_b200.layout import C, L, P, R, T, Z


def cat_pack_split(self: Union[C, P]) -> List[T]:
    data, token_sizes = self.cat()
    return torch.split(data, token_sizes.detach().cpu().tolist(), dim=0)


C.split = cat_pack_split
P.split = cat_pack_split


def _padded_split(self: Union[L, R]) -> List[T]:
    right = isinstance(self, R)
    lengths = self.token_sizes.detach().cpu().tolist()
    return tuple(row[row.size()[0] - n:] if right else row[:n] for row, n in zip(self.data.unbind(dim=0), lengths))


left_split = right_split = _padded_split
L.split = _padded_split
R.split = _padded_split


def tolist(self: Z) -> List[List[Number]]:
    return [tensor.tolist() for tensor in self.detach().split()]


C.tolist = tolist
L.tolist = tolist
P.tolist = tolist
R.tolist = tolist
