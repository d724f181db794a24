"""Reductions -- mirror of torchrua/reduce.py.

segment_* (reduce.py:34-69) run on the native segment-reduce kernel (rua_segment_reduce): one pass
over the data with fp32 accumulation, logsumexp fused, the reference's `initial` quirks reproduced
without the extra global-min pass.  scatter_* (reduce.py:6-31) are outside the hot path (unsorted
index reductions, SURVEY.md 8f "next" row 1) and still compose ATen ops exactly like the reference.
"""
import torch

from torchrua_b200 import _native
from torchrua_b200._lib import CAT, LEN_CONST, MAP_REV, MAP_SHIFT, PAD_FILL, PAD_WRAP
from torchrua_b200._native import MapSpec, SideSpec
from torchrua_b200.layout import T


def scatter_max(tensor: T, index: T, source: T, include_self: bool = False, dim: int = 0):
    return torch.index_reduce(tensor, index=index, source=source, reduce='amax', include_self=include_self, dim=dim)


def scatter_min(tensor: T, index: T, source: T, include_self: bool = False, dim: int = 0):
    return torch.index_reduce(tensor, index=index, source=source, reduce='amin', include_self=include_self, dim=dim)


def scatter_sum(tensor: T, index: T, source: T, include_self: bool = False, dim: int = 0):
    base = tensor if include_self else torch.zeros_like(tensor)
    return torch.index_add(base, index=index, source=source, dim=dim)


def scatter_mean(tensor: T, index: T, source: T, include_self: bool = False, dim: int = 0):
    return torch.index_reduce(tensor, index=index, source=source, reduce='mean', include_self=include_self, dim=dim)


def scatter_prod(tensor: T, index: T, source: T, include_self: bool = False, dim: int = 0):
    return torch.index_reduce(tensor, index=index, source=source, reduce='prod', include_self=include_self, dim=dim)


def scatter_logsumexp(tensor: T, index: T, source: T, include_self: bool = False, dim: int = 0):
    m = scatter_max(tensor, index=index, source=source, include_self=include_self, dim=dim).detach()
    shifted_self = (tensor - m).exp()
    shifted_source = (source - m[index]).exp()
    return scatter_sum(shifted_self, index=index, source=shifted_source, include_self=include_self, dim=dim).log() + m


def segment_max(tensor: T, segment_sizes: T) -> T:
    return _native.segment_reduce(tensor, segment_sizes, 'max')


def segment_min(tensor: T, segment_sizes: T) -> T:
    return _native.segment_reduce(tensor, segment_sizes, 'min')


def segment_sum(tensor: T, segment_sizes: T) -> T:
    return _native.segment_reduce(tensor, segment_sizes, 'sum')


def segment_mean(tensor: T, segment_sizes: T) -> T:
    return _native.segment_reduce(tensor, segment_sizes, 'mean')


def segment_prod(tensor: T, segment_sizes: T) -> T:
    return _native.segment_reduce(tensor, segment_sizes, 'prod')


def segment_logsumexp(tensor: T, segment_sizes: T) -> T:
    return _native.segment_reduce(tensor, segment_sizes, 'logsumexp')


def _segment_pick(tensor: T, segment_sizes: T, last: bool) -> T:
    _native.require_cuda(tensor, segment_sizes)
    rg = _native.ragged_from_lengths(segment_sizes)
    spec = MapSpec(rg=rg, src=SideSpec(CAT, rows=tensor.size()[0]),
                   dst=SideSpec(CAT, xform=LEN_CONST, arg=1, rows=rg.B),
                   tmap=MAP_REV if last else MAP_SHIFT, pad_mode=PAD_WRAP if last else PAD_FILL)
    return _native.row_map(tensor, spec)


def segment_head(tensor: T, segment_sizes: T) -> T:
    """first row of every segment (reduce.py:64-65) -- one row-map launch, B rows moved."""
    return _segment_pick(tensor, segment_sizes, last=False)


def segment_last(tensor: T, segment_sizes: T) -> T:
    """last row of every segment (reduce.py:68-69)."""
    return _segment_pick(tensor, segment_sizes, last=True)
