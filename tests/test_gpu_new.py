"""GPU parity: C/L/P/R.new(list_of_tensors) through the multi-source kernel (rua_row_map_list, host-side metadata)
against the reference's own recipe -- torch.cat, then pack_sequence / pad_sequence / F.pad (torchrua/core/__init__.py:9-36,
tests/test_layout.py of the reference builds its expectations the same way)."""
import pytest
import torch
from torch.nn import functional as F
from torch.nn.utils.rnn import pack_sequence, pad_sequence

pytestmark = pytest.mark.gpu

import torchrua_b200 as rua  # noqa: E402
from torchrua_b200 import C, L, P, R  # noqa: E402
from torchrua_b200 import core as rua_core  # noqa: E402


@pytest.fixture(params=['list_kernel', 'cat_then_convert'], autouse=True)
def both_paths(request, monkeypatch):
    """X.new picks the multi-source kernel for few large tensors and torch.cat + conversion for many small ones:
    every test runs through both."""
    monkeypatch.setattr(rua_core, 'LIST_KERNEL_MIN_BYTES_PER_TENSOR', 0 if request.param == 'list_kernel' else 1 << 60)
    return request.param


def make(lengths, feat, dtype, seed=0, requires_grad=False):
    g = torch.Generator().manual_seed(seed)
    out = []
    for n in lengths:
        t = torch.randn((n,) + feat, generator=g).mul(4)
        t = t.to(dtype).cuda()
        out.append(t.requires_grad_(True) if requires_grad else t)
    return out


CASES = [([3, 1, 5, 5, 2], (7,)), ([1], (1,)), ([4, 4, 4], ()), ([9, 2, 31, 17, 1, 1, 6], (3, 5)), (list(range(1, 70)), (64,)),
         ([2, 600, 3], (256,))]


@pytest.mark.parametrize('lengths,feat', CASES)
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16, torch.int64, torch.uint8])
def test_new_matches_cat_then_convert(lengths, feat, dtype):
    tensors = make(lengths, feat, dtype)
    lens = torch.tensor(lengths, device='cuda')
    cat = torch.cat(tensors, dim=0)
    c = C.new(tensors)
    assert torch.equal(c.data, cat) and torch.equal(c.token_sizes, lens) and c.token_sizes.dtype == torch.long
    left = L.new(tensors, 3)
    assert torch.equal(left.data, pad_sequence(tensors, batch_first=True, padding_value=3))
    assert torch.equal(left.token_sizes, lens)
    right = R.new(tensors, 3)
    t = max(lengths)
    exp = torch.stack([F.pad(x, [0, 0] * (x.dim() - 1) + [t - x.size(0), 0], value=3) for x in tensors])
    assert torch.equal(right.data, exp) and torch.equal(right.token_sizes, lens)
    pk = P.new(tensors)
    ref = pack_sequence(tensors, enforce_sorted=False)
    assert torch.equal(pk.batch_sizes, ref.batch_sizes) and not pk.batch_sizes.is_cuda
    assert torch.equal(pk.cat().data, cat)                      # canonical form (tie order is the documented deviation)
    assert torch.equal(lens[pk.sorted_indices], lens[ref.sorted_indices.to('cuda')])
    # downstream ops reuse the host-built metadata
    assert torch.equal(pk.left(0).data, L.new(tensors, 0).data)
    assert c.size()[:2] == (len(lengths), t)


def test_new_with_empty_sequences_and_views():
    base = torch.arange(200, device='cuda', dtype=torch.float32).view(20, 10)
    tensors = [base[0:3], base[3:3], base[5:9, :], base[9:10], base[10:10]]
    tensors_nc = [base[:, :6][0:3], base[:, 1:7][4:4], base[:, 2:8][5:9]]          # non-contiguous rows
    c = C.new(tensors)
    assert torch.equal(c.data, torch.cat(tensors)) and c.token_sizes.tolist() == [3, 0, 4, 1, 0]
    assert torch.equal(L.new(tensors, -1).data, pad_sequence(tensors, batch_first=True, padding_value=-1))
    assert torch.equal(C.new(tensors_nc).data, torch.cat(tensors_nc))
    pk = P.new(tensors)
    assert torch.equal(pk.cat().data, torch.cat(tensors)) and pk.cat().token_sizes.tolist() == [3, 0, 4, 1, 0]


@pytest.mark.parametrize('kind', [C, L, P, R])
def test_new_gradients_reach_every_tensor(kind):
    lengths = [3, 1, 6, 2]
    a = make(lengths, (5,), torch.float32, seed=1, requires_grad=True)
    b = [t.detach().clone().requires_grad_(True) for t in a]
    ours = kind.new(a)
    cat = C(data=torch.cat(b, dim=0), token_sizes=torch.tensor(lengths, device='cuda'))
    theirs = {C: cat.cat, L: cat.left, P: cat.pack, R: cat.right}[kind]()
    w = torch.randn_like(ours.data)
    (ours.data * w).sum().backward()
    if kind is P:   # same canonical data, possibly different tie order: weight through the cat form
        wc = P(data=w, batch_sizes=ours.batch_sizes, sorted_indices=ours.sorted_indices,
               unsorted_indices=ours.unsorted_indices).cat().data
        (theirs.cat().data * wc).sum().backward()
    else:
        (theirs.data * w).sum().backward()
    for x, y in zip(a, b):
        assert torch.equal(x.grad, y.grad)


def test_new_falls_back_to_aten_semantics_for_mixed_inputs():
    mixed = [torch.ones((2, 3), device='cuda'), torch.ones((1, 3), device='cuda', dtype=torch.float64)]
    c = C.new(mixed)                                 # type promotion like torch.cat
    assert c.data.dtype == torch.float64 and c.token_sizes.tolist() == [2, 1]
    with pytest.raises(Exception):
        C.new([])


def test_new_large_batch_device_sort_path(monkeypatch):
    from torchrua_b200 import _native
    monkeypatch.setattr(_native, 'HOST_SORT_MAX_B', 4)           # force the device-side pack metadata
    tensors = make([5, 2, 9, 1, 9, 3, 7], (4,), torch.float32)
    pk = P.new(tensors)
    ref = pack_sequence(tensors, enforce_sorted=False)
    assert torch.equal(pk.batch_sizes, ref.batch_sizes)
    assert torch.equal(pk.cat().data, torch.cat(tensors))


@pytest.mark.parametrize('lengths', [[5000, 3, 10], [4097], [4096, 4096, 1], [1, 2, 3]])
def test_pack_speculative_launch_and_its_fallback(lengths):
    """C.pack() enqueues the conversion before batch_sizes reaches the host, assuming T <= 4096 time steps; longer
    sequences invalidate the speculation and the conversion is redone with exact metadata."""
    from torchrua_b200 import _native
    _native._CACHE.clear()
    g = torch.Generator().manual_seed(3)
    tensors = [torch.randn((n, 6), generator=g).cuda().requires_grad_(True) for n in lengths]
    c = C(data=torch.cat(tensors), token_sizes=torch.tensor(lengths, device='cuda'))
    pk = c.pack()
    ref = pack_sequence(tensors, enforce_sorted=False)
    assert torch.equal(pk.batch_sizes, ref.batch_sizes)
    assert torch.equal(pk.cat().data, torch.cat(tensors))
    assert torch.equal(pk.data, ref.data) or len(set(lengths)) < len(lengths)     # identical unless ties reorder
    w = torch.randn_like(pk.data)
    (pk.data * w).sum().backward()
    got = torch.cat([t.grad for t in tensors])
    wc = P(data=w, batch_sizes=pk.batch_sizes, sorted_indices=pk.sorted_indices, unsorted_indices=pk.unsorted_indices).cat().data
    assert torch.equal(got, wc)
