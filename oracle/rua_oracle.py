"""CPU oracle for TorchRua's ragged-sequence hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the *closed forms* the reference (speedcell4/torchrua
v0.5.1, mounted read-only at /root/reference while the repo is authored) computes by composing
stock ATen ops.  It exists to check the CUDA path in ``torchrua_b200``; it is NOT part of the
product.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import it.  The product path never routes through this module and
raises when the CUDA extension is missing.

Parity status: **pinned**.  ``tests/golden/make_golden.py`` imports the live reference in the
authoring container and commits its outputs under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function below against those vectors.

Conventions: ``B`` sequences with lengths ``len[i]`` (int64), ``N = sum len``,
``T = max len``; token ``(i, t)`` lives at

    C (N, *)   : off[i] + t                       torchrua/core/get.py:25-26, layout/cat.py:79-81
    L (B,T,*)  : i*T + t                          torchrua/core/get.py:41-42, layout/left.py:73-77
    R (B,T,*)  : i*T + (T - len[i]) + t           torchrua/core/get.py:73-74, layout/right.py:74-79
    P (N, *)   : poff[t] + unsorted[i]            torchrua/core/get.py:57-58, layout/pack.py:43-45

where ``off`` / ``poff`` are exclusive prefix sums of ``len`` / ``batch_sizes``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional, Tuple

import numpy as np

I64 = np.int64


# --------------------------------------------------------------------------------------------
# containers (plain data; mirror the four reference layouts, torchrua/layout/*.py)
# --------------------------------------------------------------------------------------------
@dataclass
class Cat:  # torchrua/layout/cat.py:9-11
    data: np.ndarray
    token_sizes: np.ndarray


@dataclass
class Left:  # torchrua/layout/left.py:9-11
    data: np.ndarray
    token_sizes: np.ndarray


@dataclass
class Right:  # torchrua/layout/right.py:10-12
    data: np.ndarray
    token_sizes: np.ndarray


@dataclass
class Pack:  # torch.nn.utils.rnn.PackedSequence as patched by torchrua/layout/pack.py
    data: np.ndarray
    batch_sizes: np.ndarray
    sorted_indices: np.ndarray
    unsorted_indices: np.ndarray


# --------------------------------------------------------------------------------------------
# index primitives (torchrua/utils.py)
# --------------------------------------------------------------------------------------------
def excl_scan(sizes: np.ndarray) -> np.ndarray:
    """Exclusive prefix sum with B+1 entries (off[B] = total).  Internal helper."""
    out = np.zeros(sizes.shape[0] + 1, dtype=I64)
    np.cumsum(sizes, out=out[1:])
    return out


def get_offsets(sizes: np.ndarray) -> np.ndarray:
    """torchrua/utils.py:16-19  cumsum -> roll(1) -> [0]=0.  Raises on empty input like the reference."""
    if sizes.shape[0] == 0:
        raise IndexError('index 0 is out of bounds for dimension 0 with size 0')
    return excl_scan(np.asarray(sizes, dtype=I64))[:-1].copy()


def major_sizes_to_ptr(sizes: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """torchrua/utils.py:7-13  returns (position-within-segment, segment-id), in that order."""
    sizes = np.asarray(sizes, dtype=I64)
    off = excl_scan(sizes)
    which = np.repeat(np.arange(sizes.shape[0], dtype=I64), sizes)
    within = np.arange(off[-1], dtype=I64) - off[which]
    return within, which


def invert_permutation(perm: np.ndarray) -> np.ndarray:
    """torchrua/utils.py:22-26  out[perm[j]] = j."""
    out = np.empty_like(perm)
    out[perm] = np.arange(perm.shape[0], dtype=perm.dtype)
    return out


# --------------------------------------------------------------------------------------------
# metadata conversion (torchrua/core/view.py)
# --------------------------------------------------------------------------------------------
def pack_meta(token_sizes: np.ndarray, sorted_indices: Optional[np.ndarray] = None):
    """torchrua/core/view.py:47-58 (pack_view).

    batch_sizes[t] = #{i : len[i] > t}; sorted_indices = argsort descending.  The reference sorts
    with a NON-stable CPU torch.sort (view.py:48), so its tie order is arbitrary; pass the
    reference's ``sorted_indices`` to reproduce it bit-for-bit ("injected permutation"), else ties
    are broken by ascending index (stable) -- the documented deviation (SURVEY.md 8c hazard 1).
    """
    token_sizes = np.asarray(token_sizes, dtype=I64)
    if sorted_indices is None:
        sorted_indices = np.argsort(-token_sizes, kind='stable').astype(I64)
    unsorted_indices = invert_permutation(sorted_indices)
    t = int(token_sizes.max()) if token_sizes.shape[0] else 0
    hist = np.bincount(token_sizes, minlength=t + 1).astype(I64)
    # bs[t] = number of sequences strictly longer than t = suffix sum of the histogram
    batch_sizes = (token_sizes.shape[0] - np.cumsum(hist))[:t].astype(I64)
    return batch_sizes, sorted_indices, unsorted_indices


def lengths_from_pack(batch_sizes: np.ndarray, unsorted_indices: np.ndarray) -> np.ndarray:
    """torchrua/core/view.py:21-25 for a P source: get_mask(...).sum(1), i.e.
    len[i] = #{t : batch_sizes[t] > unsorted[i]} (batch_sizes is non-increasing)."""
    bs = np.asarray(batch_sizes, dtype=I64)
    # count of entries > r in a non-increasing array = searchsorted on the reversed (ascending) array
    asc = bs[::-1]
    return (bs.shape[0] - np.searchsorted(asc, unsorted_indices, side='right')).astype(I64)


# --------------------------------------------------------------------------------------------
# enumeration orders: ptr() of each layout (torchrua/layout/*.py)
# --------------------------------------------------------------------------------------------
def cat_ptr(token_sizes: np.ndarray):
    """torchrua/layout/cat.py:68-71 (also left.py:68-71, right.py:69-72): (batch_ptr, token_ptr)
    enumerated sequence-major."""
    token_ptr, batch_ptr = major_sizes_to_ptr(token_sizes)
    return batch_ptr, token_ptr


def pack_ptr(batch_sizes: np.ndarray, sorted_indices: np.ndarray):
    """torchrua/layout/pack.py:23-27: enumerated time-major: (sorted[r], t) for r < bs[t]."""
    rank, token_ptr = major_sizes_to_ptr(batch_sizes)
    return sorted_indices[rank], token_ptr


# --------------------------------------------------------------------------------------------
# physical row of token (i, t) in the flattened storage of each layout
# --------------------------------------------------------------------------------------------
def _rows_cat(token_sizes, b, t):
    return excl_scan(token_sizes)[b] + t


def _rows_left(token_sizes, width, b, t):
    return b * width + t


def _rows_right(token_sizes, width, b, t):
    return b * width + (width - token_sizes[b]) + t


def _rows_pack(batch_sizes, unsorted_indices, b, t):
    return excl_scan(batch_sizes)[t] + unsorted_indices[b]


def _lengths(seq) -> np.ndarray:
    if isinstance(seq, Pack):
        return lengths_from_pack(seq.batch_sizes, seq.unsorted_indices)
    return np.asarray(seq.token_sizes, dtype=I64)


def _raw(seq) -> np.ndarray:
    """raw(): torchrua/layout/cat.py:83, pack.py:51, left.py:83, right.py:85."""
    if isinstance(seq, (Left, Right)):
        return seq.data.reshape((-1,) + seq.data.shape[2:])
    return seq.data


def _rows(seq, b, t):
    if isinstance(seq, Cat):
        return _rows_cat(seq.token_sizes, b, t)
    if isinstance(seq, Left):
        return _rows_left(seq.token_sizes, seq.data.shape[1], b, t)
    if isinstance(seq, Right):
        return _rows_right(seq.token_sizes, seq.data.shape[1], b, t)
    return _rows_pack(seq.batch_sizes, seq.unsorted_indices, b, t)


# --------------------------------------------------------------------------------------------
# the 12 directed conversions (torchrua/core/cast.py)
# --------------------------------------------------------------------------------------------
def to_cat(seq) -> Cat:
    """torchrua/core/cast.py:8-16: out[off[i]+t] = X(i,t)."""
    if isinstance(seq, Cat):
        return seq
    lens = _lengths(seq)
    b, t = cat_ptr(lens)
    return Cat(data=_raw(seq)[_rows(seq, b, t)], token_sizes=lens)


def to_pack(seq, sorted_indices: Optional[np.ndarray] = None) -> Pack:
    """torchrua/core/cast.py:41-49: out[poff[t]+unsorted[i]] = X(i,t)."""
    if isinstance(seq, Pack):
        return seq
    lens = _lengths(seq)
    bs, srt, uns = pack_meta(lens, sorted_indices)
    b, t = pack_ptr(bs, srt)
    return Pack(data=_raw(seq)[_rows(seq, b, t)], batch_sizes=bs, sorted_indices=srt, unsorted_indices=uns)


def _to_padded(seq, fill_value, right: bool):
    lens = _lengths(seq)
    width = int(lens.max()) if lens.shape[0] else 0
    feat = _raw(seq).shape[1:]
    out = np.full((lens.shape[0], width) + feat, fill_value, dtype=seq.data.dtype)
    b, t = cat_ptr(lens)
    flat = out.reshape((-1,) + feat)
    dst = _rows_right(lens, width, b, t) if right else _rows_left(lens, width, b, t)
    flat[dst] = _raw(seq)[_rows(seq, b, t)]
    return out, lens


def to_left(seq, fill_value=0) -> Left:
    """torchrua/core/cast.py:19-38 (cat_pack_to_left, right_to_left): z=full(fill); z[i,t]=X(i,t)."""
    if isinstance(seq, Left):
        return seq
    out, lens = _to_padded(seq, fill_value, right=False)
    return Left(data=out, token_sizes=lens)


def to_right(seq, fill_value=0) -> Right:
    """torchrua/core/cast.py:52-71 (cat_pack_to_right, left_to_right): z[i, T-len[i]+t]=X(i,t)."""
    if isinstance(seq, Right):
        return seq
    out, lens = _to_padded(seq, fill_value, right=True)
    return Right(data=out, token_sizes=lens)


def convert(seq, kind: str, fill_value=0, sorted_indices=None):
    return {'C': to_cat, 'L': lambda s: to_left(s, fill_value), 'R': lambda s: to_right(s, fill_value),
            'P': lambda s: to_pack(s, sorted_indices)}[kind](seq)


# --------------------------------------------------------------------------------------------
# idx() / offsets() / size()  (torchrua/layout/*.py)
# --------------------------------------------------------------------------------------------
def size(seq):
    """torchrua/layout/cat.py:61-66, left.py:61-66, right.py:62-67, pack.py:12-17."""
    if isinstance(seq, Pack):
        return (int(seq.batch_sizes.max()), int(seq.batch_sizes.shape[0])) + tuple(seq.data.shape[1:])
    feat = seq.data.shape[1:] if isinstance(seq, Cat) else seq.data.shape[2:]
    return (int(seq.token_sizes.shape[0]), int(seq.token_sizes.max())) + tuple(feat)


def idx(seq):
    """cat.py:73-77 / pack.py:33-37: arange(N) in the same layout; left.py:73-77: C(i*T+t);
    right.py:74-79: C(i*T+(T-len)+t)."""
    if isinstance(seq, Cat):
        return Cat(np.arange(seq.data.shape[0], dtype=I64), seq.token_sizes)
    if isinstance(seq, Pack):
        return Pack(np.arange(seq.data.shape[0], dtype=I64), seq.batch_sizes, seq.sorted_indices,
                    seq.unsorted_indices)
    lens = np.asarray(seq.token_sizes, dtype=I64)
    width = int(lens.max())
    b, t = cat_ptr(lens)
    rows = b * width + t
    if isinstance(seq, Right):
        rows = rows + (width - lens[b])
    return Cat(rows, seq.token_sizes)


def offsets(seq):
    """cat.py:79-81: off.clamp_max(N-1); pack.py:43-45: poff.clamp_max(N-1); left.py:79-81 /
    right.py:81-83: arange(B)*T."""
    if isinstance(seq, Cat):
        return np.minimum(get_offsets(seq.token_sizes), seq.data.shape[0] - 1)
    if isinstance(seq, Pack):
        return np.minimum(get_offsets(seq.batch_sizes), seq.data.shape[0] - 1)
    lens = np.asarray(seq.token_sizes, dtype=I64)
    return np.arange(lens.shape[0], dtype=I64) * int(lens.max())


def ptr(seq):
    if isinstance(seq, Pack):
        return pack_ptr(seq.batch_sizes, seq.sorted_indices)
    return cat_ptr(seq.token_sizes)


# --------------------------------------------------------------------------------------------
# masks (torchrua/mask.py)
# --------------------------------------------------------------------------------------------
def mask(seq, zero, one, dtype) -> np.ndarray:
    """torchrua/mask.py:6-12: (B,T) of ``dtype``; ``one`` where t < len[i] else ``zero``.
    LEFT-aligned for every layout including R (mask.py:10 indexes a plain tensor with ptr())."""
    lens = _lengths(seq)
    width = int(lens.max()) if lens.shape[0] else 0
    out = np.full((lens.shape[0], width), zero, dtype=dtype)
    out[np.arange(width)[None, :] < lens[:, None]] = one
    return out


def bmask(seq) -> np.ndarray:  # torchrua/mask.py:21-22
    return mask(seq, False, True, np.bool_)


def fmask(seq, dtype=None) -> np.ndarray:  # torchrua/mask.py:31-32
    dtype = np.dtype(dtype or seq.data.dtype)
    return mask(seq, np.finfo(dtype).min, 0, dtype)


def get_mask(seq) -> np.ndarray:  # torchrua/core/view.py:11-18
    return mask(seq, 0, 1, I64)


# --------------------------------------------------------------------------------------------
# selects (torchrua/select/*.py), stated on the cat form then re-cast to the source layout
# --------------------------------------------------------------------------------------------
def _recast(cat: Cat, like, fill_value=0):
    if isinstance(like, Cat):
        return cat
    if isinstance(like, Left):
        return to_left(cat, fill_value)
    if isinstance(like, Right):
        return to_right(cat, fill_value)
    # keep the source's permutation: sorting (len - c) or equal lengths must not re-break ties
    return to_pack(cat, sorted_indices=like.sorted_indices)


def _select(seq, new_lens: np.ndarray, src_t: Callable[[np.ndarray, np.ndarray, np.ndarray], np.ndarray],
            pad_with_row0: bool = False):
    lens = _lengths(seq)
    b, t = cat_ptr(new_lens)
    rows = _rows(seq, b, src_t(b, t, lens[b]))
    out = _recast(Cat(_raw(seq)[rows], new_lens), seq)
    if pad_with_row0 and isinstance(out, (Left, Right)):
        pad = ~get_mask(out).astype(bool)
        if isinstance(out, Right):
            pad = pad[:, ::-1]
        out.data[pad] = _raw(seq)[0]
    return out


def head(seq, n: int):
    """torchrua/select/head.py:6-67: first n tokens of every sequence (requires n <= min len)."""
    lens = _lengths(seq)
    return _select(seq, np.full_like(lens, n), lambda b, t, l: t)


def last(seq) -> np.ndarray:
    """torchrua/select/last.py:7-13: X[(arange(B), len-1)] -> plain (B,*) array."""
    lens = _lengths(seq)
    b = np.arange(lens.shape[0], dtype=I64)
    if isinstance(seq, Cat):
        # cat_getitem goes through C.offsets(), which is clamped to N-1 (layout/cat.py:81); an EMPTY
        # sequence therefore reads row min(off, N-1) - 1, wrapping like any negative index
        return seq.data[np.minimum(excl_scan(lens)[:-1], seq.data.shape[0] - 1) + lens - 1]
    return _raw(seq)[_rows(seq, b, lens - 1)]


def rev(seq):
    """torchrua/select/rev.py:6-41: out(i,t) = in(i, len[i]-1-t)."""
    return _select(seq, _lengths(seq), lambda b, t, l: l - 1 - t)


def roll(seq, shifts: int):
    """torchrua/select/roll.py:6-37: out(i,t) = in(i, (t - s + len) mod len)  (floor-mod).

    Quirk (roll.py:19-20, 33-34): L/R.roll gather through an index sequence padded with index 0, so
    the padding slots of the result hold a copy of flat row 0 of the input (NOT a fill value)."""
    return _select(seq, _lengths(seq), lambda b, t, l: np.mod(t - shifts + l, l), pad_with_row0=True)


def trunc(seq, trunc_: Tuple[int, int]):
    """torchrua/select/trunc.py:9-62: out_i = in_i[a : len-b]  (requires a+b < min len)."""
    a, b_ = trunc_
    lens = _lengths(seq)
    if isinstance(seq, (Left, Right)):
        # trunc.py:26-33, 52-59: a pure slice data[:, a:T-b] -- slots outside the new lengths keep
        # whatever the input held there (stale tokens), they are NOT re-filled.
        width = int(lens.max())
        return type(seq)(seq.data[:, a:width - b_], lens - a - b_)
    return _select(seq, lens - a - b_, lambda b, t, l: t + a)


# --------------------------------------------------------------------------------------------
# segment reductions (torchrua/reduce.py:34-69  ->  ATen segment_reduce, strict left-to-right
# accumulation starting from ``initial``; SURVEY.md 8c hazards 2 and 3)
# --------------------------------------------------------------------------------------------
def _acc_dtype(dtype):
    return np.float64 if np.dtype(dtype) == np.float64 else np.float32


def _sequential(data: np.ndarray, sizes: np.ndarray, init, step):
    """acc[s] = init; for r in rows of segment s, in order: acc[s] = step(acc[s], row).
    Vectorised across segments, strictly sequential inside each one."""
    sizes = np.asarray(sizes, dtype=I64)
    off = excl_scan(sizes)
    acc = np.empty((sizes.shape[0],) + data.shape[1:], dtype=data.dtype)
    acc[...] = init
    for t in range(int(sizes.max()) if sizes.shape[0] else 0):
        live = np.nonzero(sizes > t)[0]
        acc[live] = step(acc[live], data[off[live] + t])
    return acc


def _nanmax(a, v):  # ATen: isnan(v) ? v : max(a, v) with std::max(a,b) = (a<b)?b:a
    with np.errstate(invalid='ignore'):
        return np.where(np.isnan(v), v, np.where(a < v, v, a))


def _nanmin(a, v):
    with np.errstate(invalid='ignore'):
        return np.where(np.isnan(v), v, np.where(v < a, v, a))


def segment_sum(data, sizes):
    """reduce.py:44-45  initial=0; empty segment -> 0."""
    x = data.astype(_acc_dtype(data.dtype), copy=False)
    return _sequential(x, sizes, 0, lambda a, v: a + v)


def segment_mean(data, sizes):
    """reduce.py:48-49  sum / len (len>0, non-NaN); empty segment -> 0."""
    x = data.astype(_acc_dtype(data.dtype), copy=False)
    s = _sequential(x, sizes, 0, lambda a, v: a + v)
    n = np.asarray(sizes, dtype=I64).reshape((-1,) + (1,) * (x.ndim - 1))
    with np.errstate(invalid='ignore', divide='ignore'):
        return np.where((n > 0) & ~np.isnan(s), s / np.maximum(n, 1).astype(s.dtype), s)


def segment_prod(data, sizes):
    """reduce.py:52-53  initial=1; empty segment -> 1."""
    x = data.astype(_acc_dtype(data.dtype), copy=False)
    return _sequential(x, sizes, 1, lambda a, v: a * v)


def segment_max(data, sizes):
    """reduce.py:34-36  initial = GLOBAL min of the whole tensor (a scalar over every column);
    empty segment -> that scalar; a NaN anywhere makes ``initial`` NaN and hence every output NaN."""
    x = data.astype(_acc_dtype(data.dtype), copy=False)
    init = x.min() if x.size else np.asarray(0, x.dtype)
    return _sequential(x, sizes, init, _nanmax)


def segment_min(data, sizes):
    """reduce.py:39-41  mirror of segment_max with the global max."""
    x = data.astype(_acc_dtype(data.dtype), copy=False)
    init = x.max() if x.size else np.asarray(0, x.dtype)
    return _sequential(x, sizes, init, _nanmin)


def segment_logsumexp(data, sizes):
    """reduce.py:56-61  m = segment_max; log(sum exp(x - m[seg]) + [len==0]) + m."""
    x = data.astype(_acc_dtype(data.dtype), copy=False)
    sizes = np.asarray(sizes, dtype=I64)
    m = segment_max(x, sizes)
    which = np.repeat(np.arange(sizes.shape[0], dtype=I64), sizes)
    with np.errstate(invalid='ignore', over='ignore'):
        e = np.exp(x - m[which])
    s = _sequential(e, sizes, 0, lambda a, v: a + v)
    eps = (sizes == 0).astype(x.dtype).reshape((-1,) + (1,) * (x.ndim - 1))
    with np.errstate(divide='ignore', invalid='ignore'):
        return np.log(s + eps) + m


def segment_head(data, sizes):
    """reduce.py:64-65  first row of each (non-empty) segment."""
    return data[excl_scan(np.asarray(sizes, dtype=I64))[:-1]]


def segment_last(data, sizes):
    """reduce.py:68-69  C(data, sizes).last(): last row of each segment; an EMPTY segment wraps to the
    row before its (clamped) offset -- see last()."""
    return last(Cat(data, np.asarray(sizes, dtype=I64)))


REDUCERS = {
    'sum': segment_sum, 'mean': segment_mean, 'prod': segment_prod, 'max': segment_max,
    'min': segment_min, 'logsumexp': segment_logsumexp, 'head': segment_head, 'last': segment_last,
}


# --------------------------------------------------------------------------------------------
# .seg(duration, fn)  (torchrua/segment.py)
# --------------------------------------------------------------------------------------------
def seg(seq, duration, fn):
    """segment.py:6-50.  ``duration`` is itself a ragged sequence of segment sizes summing to len[i].

    C: fn(data, duration.cat().data)                                           segment.py:6-10
    L: per row, sizes = [durations padded with 0 ..., T-len]; drop last column segment.py:16-25
    R: per row, sizes = [T-len, 0-padded durations (right aligned)]; drop col 0  segment.py:38-47
    P: via cat, then re-pack                                                    segment.py:31-32
    """
    if isinstance(seq, Cat):
        d = to_cat(duration)
        return Cat(fn(seq.data, d.data), d.token_sizes)
    if isinstance(seq, Pack):
        out = seg(to_cat(seq), duration, fn)
        return to_pack(out)
    lens = np.asarray(seq.token_sizes, dtype=I64)
    b, t = lens.shape[0], int(lens.max())
    feat = seq.data.shape[2:]
    if isinstance(seq, Left):
        d = to_left(duration, 0)
        sizes = np.concatenate([d.data, (t - lens)[:, None]], axis=-1).reshape(-1)
        out = fn(seq.data.reshape((-1,) + feat), sizes).reshape((b, -1) + feat)
        return Left(out[:, :-1], d.token_sizes)
    d = to_right(duration, 0)
    sizes = np.concatenate([(t - lens)[:, None], d.data], axis=-1).reshape(-1)
    out = fn(seq.data.reshape((-1,) + feat), sizes).reshape((b, -1) + feat)
    return Right(out[:, 1:], d.token_sizes)


# --------------------------------------------------------------------------------------------
# bf16 helpers (numpy has no bfloat16): payload moves treat bf16 as uint16; reductions follow the
# parity contract of SURVEY.md 8c hazard 2 -- upcast to fp32, reduce, round ONCE to bf16 (RNE).
# --------------------------------------------------------------------------------------------
def bf16_bits_to_f32(bits: np.ndarray) -> np.ndarray:
    return (bits.astype(np.uint32) << 16).view(np.float32)


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    rounded = (u + (np.uint32(0x7FFF) + ((u >> 16) & np.uint32(1)))) >> 16
    nan = np.isnan(x)
    out = rounded.astype(np.uint16)
    out[nan] = ((u[nan] >> 16) | np.uint32(0x0040)).astype(np.uint16)
    return out
