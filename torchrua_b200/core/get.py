"""Layout-aware indexing -- mirror of torchrua/core/get.py.  The payload rows are moved by
rua_gather_rows; only the (small) index arithmetic on user-supplied index tensors stays in torch."""
from typing import Tuple, Union

import torch
from torch import Tensor

from torchrua_b200 import _native
from torchrua_b200.layout import C, L, P, R, T, Z

Key = Union[int, Tensor, Tuple[Tensor, Tensor], Z]
Value = Union[Tensor, Z]

_SEQ = Z.__args__


def _is_pair(key) -> bool:
    return isinstance(key, tuple) and len(key) == 2 and isinstance(key[0], Tensor) and isinstance(key[1], Tensor)


def _take(rows: Tensor, index: Tensor) -> Tensor:
    """rows[index] along dim 0 through the native gather when both live on CUDA."""
    if rows.is_cuda and index.is_cuda and index.dtype in (torch.long, torch.int):
        return _native.gather_rows(rows, index)
    return super(T, rows).__getitem__(index)


def tensor_getitem(self: T, key: Key) -> Value:
    if isinstance(key, _SEQ):
        return key._replace(data=_take(self, key.data))
    return super(T, self).__getitem__(key)


T.__getitem__ = tensor_getitem


def _flat_key(self: Z, key: Tuple[Tensor, Tensor]) -> Tensor:
    """flat storage row of tokens (batch_ptr, token_ptr) -- the position table of SURVEY.md section 3."""
    b, t = key
    if isinstance(self, C):
        return self.offsets()[b] + t
    if isinstance(self, P):
        return self.unsorted_indices[b] + self.offsets()[t]
    width = self.data.size()[1]
    if isinstance(self, R):
        return b * width + (self.size()[1] - self.token_sizes[b]) + t
    return b * width + t


def sequence_getitem(self: Z, key: Key) -> Value:
    if isinstance(key, _SEQ):
        return key._replace(data=_take(self.raw(), key.data))
    if _is_pair(key):
        return _take(self.raw(), _flat_key(self, key))
    if isinstance(key, Tensor):
        return _take(self.raw(), key)
    return tuple.__getitem__(self, key)


cat_getitem = left_getitem = pack_getitem = right_getitem = sequence_getitem

C.__getitem__ = sequence_getitem
L.__getitem__ = sequence_getitem
P.__getitem__ = sequence_getitem
R.__getitem__ = sequence_getitem
