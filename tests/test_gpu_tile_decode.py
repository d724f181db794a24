"""GPU parity for the shared tile decode (csrc/tile_decode.cuh) that the narrow-row kernels, the index emitters and
the rows-on-lanes segment reduce are built on: adversarial segment structures against plain torch expressions of the
reference's closed forms (utils.py:7-13 major_sizes_to_ptr, core/cast.py, reduce.py).

Structures: thousands of EMPTY segments inside one 2048-position tile (the staged table overflows -> per-row search
fallback), segments that start / end exactly on tile boundaries, one segment covering many tiles (single-segment
fast path), runs of length-1 segments (every position starts a segment), and mixtures."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import torchrua_b200 as rua  # noqa: E402
from torchrua_b200 import C, L  # noqa: E402


def structures():
    g = torch.Generator().manual_seed(5)
    z = lambda n: torch.zeros(n, dtype=torch.long)  # noqa: E731
    o = lambda n: torch.ones(n, dtype=torch.long)   # noqa: E731
    t = lambda *v: torch.tensor(v, dtype=torch.long)  # noqa: E731
    return {
        'empties_overflow': torch.cat([t(5), z(5000), t(7, 3), z(2500), t(1)]),
        'empties_at_edges': torch.cat([z(3000), t(2048), z(2100), t(2048, 1), z(10)]),
        'tile_aligned': t(2048, 2048, 4096, 1, 2047, 2048),
        'one_huge': t(3, 50000, 2),
        'all_ones': o(9000),
        'ones_and_zeros': torch.cat([o(2047), z(1), o(2), z(4097), o(3000)]),
        'mixture': torch.cat([torch.randint(0, 3, (6000,), generator=g), t(10000), torch.randint(0, 70, (500,), generator=g)]),
        'short_tail': torch.cat([t(4096), z(2), t(1)]),
    }


CASES = structures()


@pytest.fixture(params=sorted(CASES))
def lens(request):
    return CASES[request.param]


def closed_forms(lens):
    n = int(lens.sum())
    which = torch.repeat_interleave(torch.arange(lens.numel()), lens)
    off = torch.cumsum(lens, 0) - lens
    within = torch.arange(n) - off[which]
    return n, which, within, off


def test_ptr_and_idx(lens):
    n, which, within, _ = closed_forms(lens)
    c = C(data=torch.arange(n, device='cuda'), token_sizes=lens.cuda())
    b, t = c.ptr()
    assert torch.equal(b.cpu(), which) and torch.equal(t.cpu(), within)
    width = int(lens.max())
    if lens.numel() * width <= 1 << 26:                # padded index space stays small enough to materialise
        left = L(data=torch.zeros((lens.numel(), width), device='cuda', dtype=torch.uint8), token_sizes=lens.cuda())
        assert torch.equal(left.idx().data.cpu(), which * width + within)


@pytest.mark.parametrize('dtype,feat', [(torch.int64, ()), (torch.int32, ()), (torch.float32, (3,)), (torch.bfloat16, (24,))])
def test_narrow_conversions_and_selects(lens, dtype, feat):
    n, which, within, off = closed_forms(lens)
    width = int(lens.max())
    if lens.numel() * width > 1 << 25:
        pytest.skip('padded form too large for this structure')
    g = torch.Generator().manual_seed(1)
    data = torch.randint(-99, 99, (n,) + feat, generator=g).to(dtype)
    c = C(data=data.cuda(), token_sizes=lens.cuda())
    padded = torch.zeros((lens.numel(), width) + feat, dtype=dtype)
    padded[which, within] = data
    left = c.left(0)
    assert torch.equal(left.data.cpu(), padded)                                   # C -> L   (padded kernel)
    assert torch.equal(left.cat().data.cpu(), data)                               # L -> C   (tile kernel, decoded tiles)
    right = c.right(0)
    assert torch.equal(right.cat().data.cpu(), data)                              # R -> C
    assert torch.equal(right.left(0).data.cpu(), padded)                          # R -> L
    seg_len = lens[which]
    rev_src = off[which] + (seg_len - 1 - within)
    assert torch.equal(c.rev().data.cpu(), data[rev_src])                         # C.rev   (compile-time token map)
    roll_src = off[which] + (within - 3) % seg_len
    assert torch.equal(c.roll(3).data.cpu(), data[roll_src])                      # C.roll
    if (lens > 0).all():
        pk = c.pack()
        assert torch.equal(pk.cat().data.cpu(), data)                             # C -> P -> C (transpose / tile kernels)


@pytest.mark.parametrize('dtype,hidden', [(torch.float32, 1), (torch.float32, 4), (torch.bfloat16, 1), (torch.float64, 2),
                                          (torch.bfloat16, 16), (torch.float32, 24)])
@pytest.mark.parametrize('op', ['sum', 'max', 'mean', 'logsumexp'])
def test_segment_reduce_rows_on_lanes_and_packed(lens, dtype, hidden, op):
    n = int(lens.sum())
    g = torch.Generator().manual_seed(2)
    x = torch.randn((n, hidden), generator=g).to(dtype)
    got = getattr(rua, 'segment_' + op)(x.cuda(), lens.cuda()).double().cpu()
    xd = x.double()
    lo = float(xd.min()) if n else 0.0
    rows, at = [], 0
    for k in lens.tolist():
        seg = xd[at:at + k]
        at += k
        if op == 'sum':
            rows.append(seg.sum(0))
        elif op == 'mean':
            rows.append(seg.sum(0) / k if k else torch.zeros(hidden, dtype=torch.double))
        elif op == 'max':
            rows.append(seg.max(0).values if k else torch.full((hidden,), lo, dtype=torch.double))
        else:
            rows.append(torch.logsumexp(seg, 0) if k else torch.full((hidden,), lo, dtype=torch.double))
    want = torch.stack(rows)
    tol = dict(rtol=2e-2, atol=2e-2) if dtype == torch.bfloat16 else dict(rtol=1e-5, atol=1e-4)
    if op == 'max':
        assert torch.equal(got, want)
    else:
        scale = 1.0 + float(lens.max()) ** 0.5 if dtype == torch.bfloat16 else 1.0   # bf16 sums of long segments
        torch.testing.assert_close(got, want, rtol=tol['rtol'], atol=tol['atol'] * scale)
