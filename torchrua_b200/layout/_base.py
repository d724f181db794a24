"""What the three token_sizes-carrying layouts (C, L, R) share: dtype/device casts and the cached
device metadata.  The reference spells these out per class (torchrua/layout/cat.py:13-59,
left.py:13-59, right.py:14-60); here they are generated once."""
import torch

from torchrua_b200 import _native

_DTYPE_CASTS = {
    'double': torch.double, 'float': torch.float, 'half': torch.half, 'long': torch.long,
    'int': torch.int, 'short': torch.short, 'char': torch.int8, 'byte': torch.uint8,
}


class TokenSizesOps:
    """mixin for namedtuples with fields (data, token_sizes)."""
    __slots__ = ()

    def to(self, dtype: torch.dtype = None, device: torch.device = None):
        return type(self)(
            data=self.data.to(dtype=dtype, device=device),
            token_sizes=self.token_sizes.to(device=device),
        )

    def cpu(self):
        return type(self)(data=self.data.cpu(), token_sizes=self.token_sizes.cpu())

    def cuda(self):
        return type(self)(data=self.data.cuda(), token_sizes=self.token_sizes.cuda())

    def detach(self):
        return type(self)(data=self.data.detach(), token_sizes=self.token_sizes.detach())

    def _ragged(self, want_pack: bool = False, early=None) -> '_native.Ragged':
        _native.require_cuda(self.data, self.token_sizes)
        return _native.ragged_from_lengths(self.token_sizes, want_pack, early)


def _make_cast(dtype):
    def cast(self):
        return self.to(dtype=dtype)
    return cast


for _name, _dtype in _DTYPE_CASTS.items():
    _fn = _make_cast(_dtype)
    _fn.__name__ = _name
    setattr(TokenSizesOps, _name, _fn)
