"""GPU: the maintainer's stub of INTEGRATION.md section 2, executed as written -- raw ctypes against librua_b200.so,
nothing from the torchrua_b200 Python package on the call path -- must reproduce torch's pad_sequence.  Keeps the
documented binding honest: if a signature in include/rua_b200.h moves, this breaks."""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_int32, c_int64, c_size_t, c_void_p

import pytest
import torch
from torch.nn.utils.rnn import pad_sequence

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'torchrua_b200', 'lib', 'librua_b200.so')


class Ragged(Structure):            # rua_ragged_t
    _fields_ = [('B', c_int64), ('off', c_void_p), ('poff', c_void_p),
                ('sorted', c_void_p), ('unsorted', c_void_p), ('Tp', c_int64)]


class Side(Structure):              # rua_side_t
    _fields_ = [('layout', c_int32), ('len_xform', c_int32), ('len_arg', c_int64),
                ('width', c_int64), ('rows', c_int64)]


def bind():
    lib = ctypes.CDLL(LIB)
    lib.rua_scan_workspace_bytes.restype = c_size_t
    lib.rua_scan_workspace_bytes.argtypes = [c_int64]
    lib.rua_scan_lengths.argtypes = [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.rua_row_map.argtypes = [c_void_p, c_void_p, c_int64, POINTER(Ragged), POINTER(Side), POINTER(Side),
                                c_int32, c_int64, c_int32, c_char_p, c_int32, c_void_p]
    lib.rua_error_string.restype = c_char_p
    return lib


def offsets_and_stats(lib, token_sizes):                       # replaces utils.get_offsets + size()'s max().item()
    n = token_sizes.numel()
    off = torch.empty(n + 1, dtype=torch.long, device=token_sizes.device)
    stats = torch.empty(2, dtype=torch.long, device=token_sizes.device)
    ws = torch.empty(max(lib.rua_scan_workspace_bytes(n), 8), dtype=torch.uint8, device=token_sizes.device)
    rc = lib.rua_scan_lengths(token_sizes.data_ptr(), n, 2**63 - 1, off.data_ptr(), stats.data_ptr(),
                              ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.rua_error_string(rc)
    return off, stats


def cat_pack_to_left(lib, data, token_sizes, fill_value=0):    # torchrua/core/cast.py:19-23 through the C-ABI
    off, stats = offsets_and_stats(lib, token_sizes)
    n, t = stats.tolist()
    b = token_sizes.numel()
    out = data.new_empty((b, t, *data.shape[1:]))
    rg = Ragged(b, off.data_ptr(), None, None, None, 0)
    src, dst = Side(0, 0, 0, 0, n), Side(1, 0, 0, t, b * t)            # RUA_CAT -> RUA_LEFT
    fill = bytes(torch.tensor([fill_value], dtype=data.dtype).view(torch.uint8).tolist())
    row_bytes = data[0].numel() * data.element_size()
    rc = lib.rua_row_map(data.data_ptr(), out.data_ptr(), row_bytes, rg, src, dst,
                         0, 0, 0, fill, len(fill), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.rua_error_string(rc)
    return out


@pytest.mark.parametrize('dtype,feat', [(torch.float32, (7,)), (torch.bfloat16, (64,)), (torch.int64, ())])
def test_integration_stub_matches_pad_sequence(dtype, feat):
    lib = bind()
    g = torch.Generator().manual_seed(0)
    lengths = torch.randint(1, 40, (23,), generator=g)
    seqs = [torch.randn((int(n),) + feat, generator=g).mul(9).to(dtype).cuda() for n in lengths]
    got = cat_pack_to_left(lib, torch.cat(seqs), lengths.cuda(), fill_value=3)
    assert torch.equal(got, pad_sequence(seqs, batch_first=True, padding_value=3))
