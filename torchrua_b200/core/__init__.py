"""Constructors from lists of tensors -- mirror of torchrua/core/__init__.py:9-36 (boundary only: a
torch.cat plus the host-side lengths; the conversions behind L/P/R.new are the native ones)."""
from typing import Any, List

import torch

from torchrua_b200.core.cast import *  # noqa: F401,F403
from torchrua_b200.core.get import *  # noqa: F401,F403
from torchrua_b200.core.set import *  # noqa: F401,F403
from torchrua_b200.core.view import *  # noqa: F401,F403
from torchrua_b200.layout import C, L, P, R, T


def new_cat(tensors: List[T]) -> C:
    data = torch.cat(tensors, dim=0)
    lengths = [tensor.size()[0] for tensor in tensors]
    return C(data=data, token_sizes=torch.tensor(lengths, dtype=torch.long, device=data.device))


C.new = new_cat


def new_left(tensors: List[T], fill_value: Any = 0) -> L:
    return new_cat(tensors).left(fill_value=fill_value)


L.new = new_left


def new_pack(tensors: List[T]) -> P:
    return new_cat(tensors).pack()


P.new = new_pack


def new_right(tensors: List[T], fill_value: Any = 0) -> R:
    return new_cat(tensors).right(fill_value=fill_value)


R.new = new_right
