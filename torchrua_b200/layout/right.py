"""RightAlignedSequence (R): data (B, T, *), token_sizes (B,).  Token (i, t) lives at flat row
i*T + (T - len[i]) + t.  reference: torchrua/layout/right.py:10-89."""
from collections import namedtuple

from torchrua_b200.layout._base import TokenSizesOps
from torchrua_b200.layout.left import LeftAlignedSequence


class RightAlignedSequence(TokenSizesOps, namedtuple('RightAlignedSequence', ['data', 'token_sizes'])):
    __slots__ = ()
    _right_aligned = True

    # same enumeration order and shapes as L; only the storage position of a token differs
    size = LeftAlignedSequence.size
    ptr = LeftAlignedSequence.ptr
    idx = LeftAlignedSequence.idx
    offsets = LeftAlignedSequence.offsets
    raw = LeftAlignedSequence.raw


R = RightAlignedSequence
