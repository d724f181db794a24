"""`.seg(duration, fn)` -- mirror of torchrua/segment.py: reduce each sequence over sub-segments whose
sizes are themselves a ragged sequence.  Thin dispatcher; the work is in the native conversions of
``duration`` and in one native ``fn`` call.  With one of this package's own reducers the payload is read once, from
where it lies: P.seg skips the P -> C pass and L.seg / R.seg (sum / mean / prod) skip the padding -- the reduce kernel
gathers the packed / padded storage rows in sequence order itself ("seg(mean) -> pooling")."""
import torch

from torchrua_b200 import _native, reduce as _reduce
from torchrua_b200.layout import C, L, P, R, Z

# reducers whose gathered form exists natively (rua_segment_reduce_gather): fn -> op name
_GATHERED = {
    _reduce.segment_sum: 'sum', _reduce.segment_mean: 'mean', _reduce.segment_prod: 'prod',
    _reduce.segment_max: 'max', _reduce.segment_min: 'min', _reduce.segment_logsumexp: 'logsumexp',
}


def cat_seg(self: C, duration: Z, fn) -> C:
    duration = duration.cat()
    return duration._replace(data=fn(self.data, duration.data))


C.seg = cat_seg


# reducers whose value on an EMPTY segment is a constant (the padded positions of an L / R result are empty segments of
# the reference's flattened reduction, segment.py:20,42): fn -> (op name, fill).  max / min / logsumexp put the global
# extreme of the whole padded buffer there (reduce.py:35,40,57-61) and keep the generic path.
_PADDED_FILL = {_reduce.segment_sum: ('sum', 0), _reduce.segment_mean: ('mean', 0), _reduce.segment_prod: ('prod', 1)}


def _padded_seg_gathered(self: Z, duration: Z, fn, right: bool):
    """L.seg / R.seg with segment_sum / mean / prod: the reference reduces all B x T rows of the padded buffer (padding
    included, as one extra segment per row); here the reducer gathers the N real rows through `idx()` and the (few) result
    rows are laid out left / right.  None when the generic path has to run."""
    try:
        hit = _PADDED_FILL.get(fn)
    except TypeError:
        hit = None
    if (hit is None or not self.data.is_cuda or self.data.dtype not in _native._DTYPES or _native.STRICT_REDUCTIONS
            or self.data.dim() < 2 or self.data.size()[1] != self.size()[1]):
        return None
    op, fill = hit
    duration = duration.cat()
    rows = self.idx().data                       # storage row of every real token, sequence-major
    red = _native.segment_reduce_gathered(self.raw(), rows, duration.data, op)
    out = duration._replace(data=red)
    out = out.right(fill) if right else out.left(fill)
    return out.data, out.token_sizes


def _padded_seg(self: Z, duration: Z, fn, right: bool):
    fused = _padded_seg_gathered(self, duration, fn, right)
    if fused is not None:
        return fused
    # one extra pad-segment per row soaks up the padding tokens (segment.py:20,42); its column is dropped
    b, t, *sizes = self.size()
    pad = (t - self.token_sizes)[:, None]
    if right:
        duration = duration.right(0)
        token_sizes = torch.cat([pad, duration.data], dim=-1).view(-1)
    else:
        duration = duration.left(0)
        token_sizes = torch.cat([duration.data, pad], dim=-1).view(-1)
    data = fn(self.data.flatten(start_dim=0, end_dim=1), token_sizes).view((b, -1, *sizes))
    return (data[:, 1:] if right else data[:, :-1]), duration.token_sizes


def left_seg(self: L, duration: Z, fn) -> L:
    data, token_sizes = _padded_seg(self, duration, fn, right=False)
    return L(data=data, token_sizes=token_sizes)


L.seg = left_seg


def pack_seg(self: P, duration: Z, fn) -> P:
    # reference (segment.py:32-33): P -> C (2 N D bytes), reduce, C -> P.  The time-major rows never need to be
    # materialised sequence-major: reduce over `P.idx()` in sequence order (N int64 instead of the N D payload pass).
    try:
        op = _GATHERED.get(fn)
    except TypeError:               # an unhashable callable
        op = None
    if op is not None and self.data.is_cuda and self.data.dtype in _native._DTYPES and not _native.STRICT_REDUCTIONS:
        duration = duration.cat()
        rows = self.idx().cat().data
        return duration._replace(data=_native.segment_reduce_gathered(self.data, rows, duration.data, op)).pack()
    return self.cat().seg(duration, fn).pack()


P.seg = pack_seg


def right_seg(self: R, duration: Z, fn) -> R:
    data, token_sizes = _padded_seg(self, duration, fn, right=True)
    return R(data=data, token_sizes=token_sizes)


R.seg = right_seg
