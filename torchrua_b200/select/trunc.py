"""trunc((a, b)): drop a leading and b trailing tokens of every sequence (requires a+b < min length)
-- mirror of torchrua/select/trunc.py.  L/R are slices like the reference (trunc.py:26-33,52-59);
C and P are one row-map launch with the new offsets in closed form (off[i] - i*(a+b), poff[t+a+b])."""
from typing import Tuple

from torchrua_b200 import _native
from torchrua_b200._lib import CAT, LEN_MINUS, MAP_SHIFT, PACK
from torchrua_b200._native import MapSpec, SideSpec
from torchrua_b200.core.cast import side_of
from torchrua_b200.layout import C, L, P, R


def cat_trunc(self: C, trunc: Tuple[int, int]) -> C:
    a, b = trunc
    rg = self._ragged()
    rows = self.data.size()[0] - rg.B * (a + b)
    spec = MapSpec(rg=rg, src=side_of(self, rg), dst=SideSpec(CAT, xform=LEN_MINUS, arg=a + b, rows=rows),
                   tmap=MAP_SHIFT, tmap_arg=a)
    return C(data=_native.row_map(self.data, spec), token_sizes=self.token_sizes - a - b)


C.trunc = cat_trunc


def _padded_trunc(self, trunc: Tuple[int, int]):
    a, b = trunc
    _, t, *_ = self.size()
    return type(self)(data=self.data[:, a:t - b], token_sizes=self.token_sizes - a - b)


left_trunc = right_trunc = _padded_trunc
L.trunc = _padded_trunc
R.trunc = _padded_trunc


def pack_trunc(self: P, trunc: Tuple[int, int]) -> P:
    a, b = trunc
    rg = self._ragged()
    batch_sizes = self.batch_sizes[a + b:]
    rows = int(batch_sizes.sum())
    spec = MapSpec(rg=rg, src=side_of(self, rg), dst=SideSpec(PACK, xform=LEN_MINUS, arg=a + b, rows=rows),
                   tmap=MAP_SHIFT, tmap_arg=a)
    return self._replace(data=_native.row_map(self.data, spec), batch_sizes=batch_sizes)


P.trunc = pack_trunc
