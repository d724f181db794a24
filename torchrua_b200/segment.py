"""`.seg(duration, fn)` -- mirror of torchrua/segment.py: reduce each sequence over sub-segments whose
sizes are themselves a ragged sequence.  Thin dispatcher; the work is in the native conversions of
``duration`` and in one native ``fn`` call."""
import torch

from torchrua_b200.layout import C, L, P, R, Z


def cat_seg(self: C, duration: Z, fn) -> C:
    duration = duration.cat()
    return duration._replace(data=fn(self.data, duration.data))


C.seg = cat_seg


def _padded_seg(self: Z, duration: Z, fn, right: bool):
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
    return self.cat().seg(duration, fn).pack()


P.seg = pack_seg


def right_seg(self: R, duration: Z, fn) -> R:
    data, token_sizes = _padded_seg(self, duration, fn, right=True)
    return R(data=data, token_sizes=token_sizes)


R.seg = right_seg
