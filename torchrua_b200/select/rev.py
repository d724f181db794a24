"""rev(): reverse every sequence in place of its layout -- mirror of torchrua/select/rev.py.
out(i, t) = in(i, len[i]-1-t); one launch for every layout (the reference's P.rev is 404 ATen calls)."""
from torchrua_b200._lib import MAP_REV
from torchrua_b200.layout import C, L, P, R, Z
from torchrua_b200.select._common import same_layout_map


def cat_rev(self: Z) -> Z:
    return same_layout_map(self, MAP_REV)


left_rev = pack_rev = right_rev = cat_rev   # one kernel serves all four layouts

for _cls in (C, L, P, R):
    _cls.rev = cat_rev
