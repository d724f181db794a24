"""Constructors from lists of tensors -- mirror of torchrua/core/__init__.py:9-36.  Boundary only: one
torch.cat plus the host-side lengths; everything after that (L/P/R.new) is a native conversion."""
from typing import Any, List

import torch

from torchrua_b200.core.cast import *  # noqa: F401,F403
from torchrua_b200.core.get import *  # noqa: F401,F403
from torchrua_b200.core.set import *  # noqa: F401,F403
from torchrua_b200.core.view import *  # noqa: F401,F403
from torchrua_b200.layout import C, L, P, R, T


def new_cat(tensors: List[T]) -> C:
    """concatenate along dim 0 and remember where each tensor ended."""
    lengths = torch.tensor([t.size()[0] for t in tensors], dtype=torch.long)
    data = torch.cat(tensors, dim=0)
    return C(data=data, token_sizes=lengths.to(device=data.device))


def new_pack(tensors: List[T]) -> P:
    return new_cat(tensors).pack()


def new_left(tensors: List[T], fill_value: Any = 0) -> L:
    return new_cat(tensors).left(fill_value=fill_value)


def new_right(tensors: List[T], fill_value: Any = 0) -> R:
    return new_cat(tensors).right(fill_value=fill_value)


for _cls, _ctor in ((C, new_cat), (L, new_left), (P, new_pack), (R, new_right)):
    _cls.new = _ctor
