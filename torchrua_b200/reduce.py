"""Reductions -- mirror of torchrua/reduce.py.

segment_* (reduce.py:34-69) run on the native segment-reduce kernel (rua_segment_reduce): one pass
over the data with fp32 accumulation, logsumexp fused, the reference's `initial` quirks reproduced
without the extra global-min pass.  scatter_* (reduce.py:6-31; SURVEY.md 8f "next" row 1) sort the
index on the device and run the same kernel over gathered rows instead of ATen's atomics.
"""
import torch

from torchrua_b200 import _native
from torchrua_b200._lib import CAT, LEN_CONST, MAP_REV, MAP_SHIFT, PAD_FILL, PAD_WRAP
from torchrua_b200._native import MapSpec, SideSpec, strict_reductions  # noqa: F401  (extension: parity mode)
from torchrua_b200.layout import T


_FLOATS = (torch.float16, torch.bfloat16, torch.float32, torch.float64)


def _scatter(tensor: T, index: T, source: T, include_self: bool, dim: int, op: str) -> T:
    """out[m] = op over {source[k] : index[k] == m} (and tensor[m] if include_self); rows no index points at
    keep tensor[m] (index_reduce semantics, reduce.py:6-23) -- except for `sum`, whose reference goes through
    index_add on zeros (reduce.py:14-15).

    Native path: stable device sort of `index`, bucket boundaries, then ONE segment-reduce launch that gathers
    the source rows in sorted order inside the kernel (rua_segment_reduce_gather).  Deterministic: no atomics.
    The O(M) combination with `tensor` is elementwise torch code, which also gives its gradient."""
    if dim != 0:
        out = _scatter(tensor.movedim(dim, 0), index, source.movedim(dim, 0), include_self, 0, op)
        return out.movedim(0, dim)
    native = 'sum' if op == 'mean' else op
    reduced, count = _native.scatter_reduce(source, index, tensor.size()[0], native)
    shape = (-1,) + (1,) * (tensor.dim() - 1)
    touched = (count > 0).view(shape)
    if op == 'sum':
        return tensor + reduced if include_self else reduced
    if op == 'mean':
        denom = (count + int(include_self)).clamp_min(1).to(dtype=tensor.dtype).view(shape)
        total = reduced + tensor if include_self else reduced
        return torch.where(touched, total / denom, tensor)
    if op == 'prod':
        return torch.where(touched, reduced * tensor if include_self else reduced, tensor)
    if op == 'max':
        return torch.where(touched, torch.maximum(reduced, tensor) if include_self else reduced, tensor)
    if op == 'min':
        return torch.where(touched, torch.minimum(reduced, tensor) if include_self else reduced, tensor)
    # logsumexp (reduce.py:26-31): untouched rows give log(exp(0)) + tensor with include_self, log(0) without
    if include_self:
        return torch.where(touched, torch.logaddexp(reduced, tensor), tensor)
    return torch.where(touched, reduced, torch.full_like(reduced, float('-inf')))


def _scatter_aten(tensor, index, source, include_self, dim, reduce):
    # integer payloads: the native reduction kernels are floating point only (like segment_reduce)
    if reduce == 'sum':
        return torch.index_add(tensor if include_self else torch.zeros_like(tensor), index=index, source=source, dim=dim)
    return torch.index_reduce(tensor, index=index, source=source, reduce=reduce, include_self=include_self, dim=dim)


def _scatter_dispatch(op, aten_name):
    def scatter(tensor: T, index: T, source: T, include_self: bool = False, dim: int = 0):
        if tensor.dtype in _FLOATS and source.dtype == tensor.dtype:
            return _scatter(tensor, index, source, include_self, dim, op)
        return _scatter_aten(tensor, index, source, include_self, dim, aten_name)
    scatter.__name__ = 'scatter_' + op
    return scatter


scatter_max = _scatter_dispatch('max', 'amax')
scatter_min = _scatter_dispatch('min', 'amin')
scatter_sum = _scatter_dispatch('sum', 'sum')
scatter_mean = _scatter_dispatch('mean', 'mean')
scatter_prod = _scatter_dispatch('prod', 'prod')
scatter_logsumexp = _scatter_dispatch('logsumexp', None)


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
