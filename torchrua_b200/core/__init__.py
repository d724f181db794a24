"""Constructors from lists of tensors -- mirror of torchrua/core/__init__.py:9-36.

The reference concatenates the list (torch.cat: every byte read and written once) and then converts (every byte
moved again).  Here the lengths are host-side shapes, so ALL metadata (offsets, N, T, the stable descending order,
batch_sizes) is computed on the host and uploaded in one copy -- no device sync -- and one multi-source kernel
(rua_row_map_list) reads every sequence from its own allocation straight into the target layout."""
from typing import Any, List

import torch

from torchrua_b200.core.cast import *  # noqa: F401,F403
from torchrua_b200.core.get import *  # noqa: F401,F403
from torchrua_b200.core.set import *  # noqa: F401,F403
from torchrua_b200.core.view import *  # noqa: F401,F403
from torchrua_b200 import _native
from torchrua_b200._lib import CAT, LEFT, PACK, RIGHT
from torchrua_b200.layout import C, L, P, R, T


def new_cat(tensors: List[T]) -> C:
    """concatenate along dim 0 and remember where each tensor ended (core/__init__.py:9-12).  The copy is
    torch.cat's (nothing to fuse: every byte moves once either way); the lengths are host-side shapes, so the
    offsets / N / T go up with them in one copy and no later op has to read them back from the device."""
    data = torch.cat(tensors, dim=0)
    lengths = [t.size()[0] for t in tensors]
    if data.is_cuda:
        rg, _ = _native.ragged_from_host_lengths(lengths, data.device, want_pack=False)
        return C(data=data, token_sizes=rg.len)
    return C(data=data, token_sizes=torch.tensor(lengths, dtype=torch.long))


# The multi-source kernel saves one full pass over the payload (2 * bytes / 6.4 TB/s) but its host side walks the
# list in Python (~0.6 us per tensor more than torch.cat's C++ loop): it pays off from ~2 MB per tensor.  Measured
# on a B200 at 4096 tensors x 0.5 MB: list kernel 3.7 ms vs torch.cat + conversion 2.7 ms (host-bound).
LIST_KERNEL_MIN_BYTES_PER_TENSOR = 2 << 20


def _new(tensors: List[T], layout: int, fill_value: Any = 0):
    """L / P / R.new.  Few large tensors: one multi-source kernel straight into the target layout (no intermediate
    concatenation).  Many small tensors, or lists the kernel cannot take (mixed dtypes / devices, CPU): torch.cat,
    then one conversion -- either way the metadata is built on the host from the shapes: no device sync."""
    if len(tensors) > 0 and isinstance(tensors[0], torch.Tensor) and tensors[0].is_cuda:
        nbytes = sum(map(torch.Tensor.numel, tensors)) * tensors[0].element_size()
        if nbytes >= LIST_KERNEL_MIN_BYTES_PER_TENSOR * len(tensors):
            plan = _native.list_plan(tensors)
            if plan is not None:
                return _native.new_from_list(tensors, layout, fill_value, plan)
    return None, None


def new_pack(tensors: List[T]) -> P:
    data, rg = _new(tensors, PACK)
    if rg is None:
        return new_cat(tensors).pack()
    return P(data=data, batch_sizes=rg.bs_cpu, sorted_indices=rg.sorted, unsorted_indices=rg.unsorted)


def new_left(tensors: List[T], fill_value: Any = 0) -> L:
    data, rg = _new(tensors, LEFT, fill_value)
    if rg is None:
        return new_cat(tensors).left(fill_value=fill_value)
    return L(data=data, token_sizes=rg.len)


def new_right(tensors: List[T], fill_value: Any = 0) -> R:
    data, rg = _new(tensors, RIGHT, fill_value)
    if rg is None:
        return new_cat(tensors).right(fill_value=fill_value)
    return R(data=data, token_sizes=rg.len)


for _cls, _ctor in ((C, new_cat), (L, new_left), (P, new_pack), (R, new_right)):
    _cls.new = _ctor
