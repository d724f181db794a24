"""Layout-aware indexing -- mirror of torchrua/core/get.py.  Position keys (batch_ptr, token_ptr) become flat storage
rows in one small kernel (rua_token_rows: offsets clamp, right-alignment shift, negative-index wrap, bounds check) and
the payload rows are moved by rua_gather_rows; gradients flow back through a native sort + segment-sum."""
from typing import Tuple, Union

import torch
from torch import Tensor

from torchrua_b200 import _native
from torchrua_b200.core.cast import side_of
from torchrua_b200.layout import C, L, P, R, T, Z

Key = Union[int, Tensor, Tuple[Tensor, Tensor], Z]
Value = Union[Tensor, Z]

_SEQ = Z.__args__
_INDEX_DTYPES = (torch.long, torch.int)


def _is_pair(key) -> bool:
    return isinstance(key, tuple) and len(key) == 2 and isinstance(key[0], Tensor) and isinstance(key[1], Tensor)


def _native_index(rows: Tensor, index: Tensor) -> bool:
    return rows.is_cuda and index.is_cuda and index.dtype in _INDEX_DTYPES


def _take(rows: Tensor, index: Tensor) -> Tensor:
    """rows[index] along dim 0 through the native gather when both live on CUDA (integer index tensors; boolean
    masks and CPU tensors keep ATen's semantics and error messages)."""
    if _native_index(rows, index):
        return _native.gather_rows(rows, index)
    return super(T, rows).__getitem__(index)


def tensor_getitem(self: T, key: Key) -> Value:
    if isinstance(key, _SEQ):
        return key._replace(data=_take(self, key.data))
    return super(T, self).__getitem__(key)


T.__getitem__ = tensor_getitem


def _flat_key(self: Z, key: Tuple[Tensor, Tensor]) -> Tensor:
    """flat storage row of tokens (batch_ptr, token_ptr) -- get.py:25-26 (C), :42 (L), :58 (P), :74 (R)."""
    b, t = key
    if not (b.is_cuda and t.is_cuda and b.dtype in _INDEX_DTYPES and t.dtype in _INDEX_DTYPES and self.data.is_cuda):
        raise RuntimeError('torchrua_b200: (batch_ptr, token_ptr) keys must be integer CUDA tensors (no CPU fallback)')
    rg = self._ragged()
    big_t = rg.T if isinstance(self, R) else 0      # R shifts by size()[1] = max length, not by the storage width
    return _native.token_rows(rg, side_of(self, rg), big_t, b, t)


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
