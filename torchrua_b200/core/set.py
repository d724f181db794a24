"""Layout-aware in-place assignment -- mirror of torchrua/core/set.py (rua_scatter_rows moves the rows)."""
from typing import Tuple, Union

import torch
from torch import Tensor

from torchrua_b200 import _native
from torchrua_b200.core.get import _SEQ, _flat_key, _is_pair
from torchrua_b200.layout import C, L, P, R, T, Z

Key = Union[int, Tensor, Tuple[Tensor, Tensor], Z]


def _put(rows: Tensor, index: Tensor, value) -> None:
    tracked = rows.requires_grad or (isinstance(value, Tensor) and value.requires_grad)
    if rows.is_cuda and index.is_cuda and index.dtype in (torch.long, torch.int) and rows.is_contiguous() \
            and not (tracked and torch.is_grad_enabled()):
        _native.scatter_rows_(rows, index, value)
    else:
        super(T, rows).__setitem__(index, value)


def tensor_setitem(self: T, key: Key, value: Tensor) -> None:
    if isinstance(key, _SEQ):
        _put(self, key.data, value)
        return None
    return super(T, self).__setitem__(key, value)


T.__setitem__ = tensor_setitem


def sequence_setitem(self: Z, key: Key, value: Tensor) -> None:
    if isinstance(key, _SEQ):
        _put(self.raw(), key.data, value)
        return None
    if _is_pair(key):
        _put(self.raw(), _flat_key(self, key), value)
        return None
    if isinstance(key, Tensor):
        _put(self.raw(), key, value)
        return None
    raise TypeError(f'{type(self).__name__} does not support item assignment with key {type(key).__name__}')


cat_setitem = left_setitem = pack_setitem = right_setitem = sequence_setitem

C.__setitem__ = sequence_setitem
L.__setitem__ = sequence_setitem
P.__setitem__ = sequence_setitem
R.__setitem__ = sequence_setitem
