"""Layout-aware in-place assignment -- mirror of torchrua/core/set.py.  rua_scatter_rows moves the rows; under
autograd the same kernel runs inside an in-place autograd function (the reference's IndexPutBackward0)."""
from typing import Tuple, Union

import torch
from torch import Tensor

from torchrua_b200 import _native
from torchrua_b200.core.get import _SEQ, _flat_key, _is_pair, _native_index
from torchrua_b200.layout import C, L, P, R, T, Z

Key = Union[int, Tensor, Tuple[Tensor, Tensor], Z]


def _put(rows: Tensor, index: Tensor, value) -> None:
    if _native_index(rows, index) and rows.is_contiguous():
        tracked = torch.is_grad_enabled() and (rows.requires_grad or (isinstance(value, Tensor) and value.requires_grad))
        if tracked:
            _native.scatter_rows_tracked_(rows, index, value)
        else:
            _native.scatter_rows_(rows, index, value)
    else:   # boolean masks, CPU tensors, strided destinations: ATen's semantics and error messages
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
        if isinstance(self, (L, R)) and not self.data.is_contiguous():
            # raw() of a strided padded tensor is a COPY; the reference writes through self.data[b, col] (set.py:47,87)
            b, t = key
            col = t if isinstance(self, L) else self.size()[1] - self.token_sizes[b] + t
            super(T, self.data).__setitem__((b, col), value)
            return None
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
