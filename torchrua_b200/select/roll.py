"""roll(shifts): circular shift inside every sequence -- mirror of torchrua/select/roll.py.
out(i, t) = in(i, (t - shifts) mod len[i]) with a non-negative (floor) modulus, any integer shift."""
from torchrua_b200._lib import MAP_ROLL, PAD_FILL, PAD_ROW0
from torchrua_b200.layout import C, L, P, R, Z
from torchrua_b200.select._common import same_layout_map


def cat_roll(self: Z, shifts: int) -> Z:
    # L/R: the reference gathers through an index sequence padded with index 0, so padding slots of the
    # result hold flat row 0 of the input (roll.py:19-20,33-34); reproduced for bit-exactness
    padded = isinstance(self, (L, R))
    return same_layout_map(self, MAP_ROLL, int(shifts), PAD_ROW0 if padded else PAD_FILL)


left_roll = pack_roll = right_roll = cat_roll   # one kernel serves all four layouts

for _cls in (C, L, P, R):
    _cls.roll = cat_roll
