"""shared plumbing of the selects: describe (source side, destination side, token map) and launch the
row-map kernel once."""
from torchrua_b200 import _native
from torchrua_b200._lib import CAT, LEFT, LEN_CONST, LEN_MINUS, LEN_SAME, PACK, RIGHT
from torchrua_b200._native import MapSpec, SideSpec
from torchrua_b200.core.cast import side_of
from torchrua_b200.layout import C, L, P, R, Z


def same_layout_map(self: Z, tmap: int, tmap_arg: int = 0, pad_mode: int = 0):
    """out has the layout, lengths and shape of ``self``; only the token order changes (rev / roll)."""
    rg = self._ragged()
    side = side_of(self, rg)
    spec = MapSpec(rg=rg, src=side, dst=side, tmap=tmap, tmap_arg=tmap_arg, pad_mode=pad_mode)
    out = _native.row_map(self.raw(), spec)
    return self._replace(data=out.view(self.data.size()))
