"""last(): the final token of every sequence as a plain (B, *) tensor -- mirror of
torchrua/select/last.py:7-13; one row-map launch moving B rows (no cat_view mask detour)."""
from torch import Tensor

from torchrua_b200 import _native
from torchrua_b200._lib import CAT, LEN_CONST, MAP_REV, PAD_WRAP
from torchrua_b200._native import MapSpec, SideSpec
from torchrua_b200.core.cast import side_of
from torchrua_b200.layout import C, L, P, R, Z


def last(self: Z) -> Tensor:
    rg = self._ragged()
    spec = MapSpec(rg=rg, src=side_of(self, rg), dst=SideSpec(CAT, xform=LEN_CONST, arg=1, rows=rg.B), tmap=MAP_REV,
                   pad_mode=PAD_WRAP)
    return _native.row_map(self.raw(), spec)


C.last = last
L.last = last
P.last = last
R.last = last
