"""Index primitives -- mirror of torchrua/utils.py, bodies replaced by the K0/K3 kernels.

reference: get_offsets utils.py:16-19, major_sizes_to_ptr :7-13, invert_permutation :22-26,
to_self :29, with_shape/broadcast_*/gather :33-51 (host-side shape helpers, unchanged semantics).
"""
from typing import Any, List, Tuple

import torch
from torch import Tensor

from torchrua_b200 import _native


def get_offsets(sizes: Tensor) -> Tensor:
    """exclusive prefix sum of ``sizes`` (one chained-scan kernel instead of cumsum/roll/index_put)."""
    if sizes.size()[0] == 0:  # the reference fails here too (utils.py:18 writes sizes[0])
        raise IndexError('index 0 is out of bounds for dimension 0 with size 0')
    _native.require_cuda(sizes)
    off, _ = _native.scan(sizes)
    return off[:-1]


def major_sizes_to_ptr(sizes: Tensor) -> Tuple[Tensor, Tensor]:
    """(position within segment, segment id) for every element of the segmented range, in that order
    (utils.py:7-13).  One emit kernel; the total is read from the scan kernel's pinned-memory notice."""
    _native.require_cuda(sizes)
    off, n, _ = _native.scan_with_totals(sizes)
    which, within, _ = _native.emit_ptr(off, n)
    return within, which


def invert_permutation(tensor: Tensor) -> Tensor:
    return _native.invert_permutation(tensor)


def to_self(self: Any, *_, **__) -> Any:
    return self


# The four helpers below are NOT on the hot path (no other file of the reference uses them); they are
# kept, with the reference's semantics (utils.py:33-51), only so that `from torchrua import *` exposes the same names --
# tests/test_abi.py compares the public surface name by name.  Host-side shape arithmetic over torch views: no kernel.
def with_shape(shape: torch.Size, dim: int, value: int) -> List[int]:
    out = list(shape)
    out[dim] = value
    return out


def broadcast_shapes(*sizes: torch.Size, dim: int):
    common = torch.broadcast_shapes(*(with_shape(s, dim=dim, value=1) for s in sizes))
    return [with_shape(common, dim=dim, value=s[dim]) for s in sizes]


def broadcast_tensors(*tensors: Tensor, dim: int):
    shapes = broadcast_shapes(*(t.size() for t in tensors), dim=dim)
    return [t.expand(shape) for t, shape in zip(tensors, shapes)]


def gather(tensor: Tensor, index: Tensor, dim: int) -> Tensor:
    tensor, index = broadcast_tensors(tensor, index, dim=dim)
    return tensor.gather(dim=dim, index=index)
