from typing import Union

from torch import Tensor

from torchrua_b200.layout.cat import C, CattedSequence
from torchrua_b200.layout.left import L, LeftAlignedSequence
from torchrua_b200.layout.pack import P, PackedSequence, idx, offsets, ptr, raw, size  # free functions leak in the reference too
from torchrua_b200.layout.right import R, RightAlignedSequence
from torchrua_b200.utils import get_offsets, major_sizes_to_ptr

T = Tensor
Z = Union[C, L, P, R]
