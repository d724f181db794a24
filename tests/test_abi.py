"""CPU-only checks of the boundary: the C-ABI library loads without a GPU and exports every symbol
include/rua_b200.h declares; the Python mirror exposes the reference's public names; the product path
refuses CPU tensors loudly instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as entry
    entry.build_cuda()
    from torchrua_b200 import _lib
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'rua_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(rua_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_exported(lib):
    from torchrua_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), f'{name} declared in include/rua_b200.h but not exported'
        assert name in _lib.SIGNATURES, f'{name} has no ctypes signature in torchrua_b200/_lib.py'
    assert sorted(_lib.SIGNATURES) == names


def test_no_torch_types_cross_the_abi():
    text = open(os.path.join(ROOT, 'include', 'rua_b200.h')).read()
    assert 'torch' not in re.sub(r'/\*.*?\*/', '', text, flags=re.S).lower()
    assert 'extern "C"' in text


def test_host_only_entry_points(lib):
    assert lib.rua_version() >= 100
    assert lib.rua_error_string(0) == b'ok'
    assert lib.rua_error_string(-2) == b'workspace too small'
    assert lib.rua_scan_workspace_bytes(1) >= 8
    assert lib.rua_scan_workspace_bytes(1 << 20) >= (1 << 20) // 4096 * 8
    assert lib.rua_sort_workspace_bytes(1 << 20) >= 4 * 4 * (1 << 20)
    # argument validation happens before any CUDA call: safe without a GPU
    assert lib.rua_scan_lengths(None, -1, 0, None, None, None, 0, None) == -1
    assert lib.rua_segment_reduce(None, None, -1, 0, 0, 0, 0, None, None, 0, None) == -1
    assert lib.rua_mask(None, -1, 0, None, None, 1, None, None) == -1
    # the entry points added for parity mode, list constructors and the multi-GPU gather validate the same way
    assert lib.rua_segment_reduce_strict(None, None, -1, 0, 0, 0, 0, None, None) == -1
    assert lib.rua_segment_reduce_strict(None, None, 4, 2, 8, 0, 3, None, None) == -1          # null buffers
    assert lib.rua_scatter_rows_multi(None, None, -1, 8, None, 1, None) == -1
    assert lib.rua_scatter_rows_multi(None, None, 4, 8, None, 99, None) == -1                   # too many destinations
    assert lib.rua_row_map_multi(None, 8, None, None, 4, None, None, 1, None) == -1
    assert lib.rua_row_map_list(None, 16, None, 8, None, None, None, 0, None) == -1
    assert lib.rua_peer_window_alloc(0, None, None) == -1
    assert lib.rua_peer_window_open(None, None) == -1
    assert lib.rua_launch_count() == 0


def test_host_side_selftest(lib):
    """launch-time arithmetic of the narrow-row kernels (magic-number division, exact for every x < 2^31): checked on
    the host by the library itself, no GPU involved."""
    assert lib.rua_selftest() == 0


def test_api_surface_matches_reference():
    """names recorded from dir(torchrua) of the live reference by tests/golden/make_golden.py"""
    import torchrua_b200 as rua
    names = [n for n in open(os.path.join(ROOT, 'tests', 'golden', 'api_names.txt')).read().split() if n]
    missing = [n for n in names if not hasattr(rua, n)]
    assert not missing, f'missing public names: {missing}'
    methods = [line.split() for line in open(os.path.join(ROOT, 'tests', 'golden', 'api_methods.txt')) if line.strip()]
    kinds = {'C': rua.C, 'L': rua.L, 'P': rua.P, 'R': rua.R}
    gaps = [(k, m) for k, m in methods if not hasattr(kinds[k], m)]
    assert not gaps, f'missing methods: {gaps}'


def test_cpu_tensors_are_rejected_loudly():
    import torchrua_b200 as rua
    c = rua.C(data=torch.randn(5, 2), token_sizes=torch.tensor([2, 3]))
    for call in (lambda: c.pack(), lambda: c.left(), lambda: c.right(), lambda: c.bmask(), lambda: c.rev(),
                 lambda: c.last(), lambda: c.ptr(), lambda: c.offsets(), lambda: c.size(),
                 lambda: rua.segment_sum(c.data, c.token_sizes), lambda: rua.segment_logsumexp(c.data, c.token_sizes),
                 lambda: rua.get_offsets(c.token_sizes)):
        with pytest.raises(RuntimeError, match='CUDA tensors only'):
            call()


def test_missing_library_raises(monkeypatch):
    from torchrua_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', '/nonexistent/librua_b200.so')
    with pytest.raises(RuntimeError, match='no CPU / PyTorch fallback'):
        _lib.load()


def test_tensor_indexing_still_works_after_patch():
    import torchrua_b200  # noqa: F401  (patches Tensor.__getitem__/__setitem__ like the reference)
    x = torch.arange(12).view(3, 4)
    assert x[1, 2].item() == 6
    assert x[torch.tensor([0, 2])].shape == (2, 4)
    x[0] = 7
    assert int(x[0].sum()) == 28
    c = torchrua_b200.C(data=torch.tensor([2, 0, 1]), token_sizes=torch.tensor([1, 2]))
    picked = torch.arange(10, 13)[c]       # Tensor[Z] -> Z with gathered data (core/get.py:11-15)
    assert isinstance(picked, torchrua_b200.C) and picked.data.tolist() == [12, 10, 11]
    data, sizes = c                         # namedtuple unpacking
    assert c[0] is data and c[1] is sizes


def test_seg_fast_paths_recognise_the_public_reducers():
    """`.seg(duration, fn)` picks its gathered / short-segment paths by the identity of `fn`: the reducers exported by the
    package AND by the alias package (`import torchrua`) must be the very functions the dispatch tables hold."""
    import torchrua
    import torchrua_b200
    from torchrua_b200 import segment
    for name in ('sum', 'mean', 'prod', 'max', 'min', 'logsumexp'):
        fn = getattr(torchrua_b200, 'segment_' + name)
        assert getattr(torchrua, 'segment_' + name) is fn
        assert segment._GATHERED[fn] == name
    assert {v[0] for v in segment._PADDED_FILL.values()} == {'sum', 'mean', 'prod'}
    assert segment._GATHERED.get(lambda t, s: t) is None
