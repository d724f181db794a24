"""GPU parity of the warp-per-32-segments reduction (csrc/reduce_warpseg.cu): narrow rows (<= 16 bytes) and at least
32768 segments -- the BASELINE config-5 shape (per-token scalars reduced per sequence).  Against the numpy oracle:
max / min bit-exact, sum / mean / logsumexp / prod within the north_star tolerances."""
import numpy as np
import pytest
import torch

from oracle import rua_oracle as ora

pytestmark = pytest.mark.gpu

S_MIN = 32768


@pytest.fixture(scope='module')
def rua():
    import torchrua_b200
    return torchrua_b200


def host(t):
    t = t.detach().cpu().contiguous()
    return t.view(torch.uint16).numpy() if t.dtype == torch.bfloat16 else t.numpy()


def to_f32(a, dtype):
    if dtype == torch.bfloat16:
        return ora.bf16_bits_to_f32(a)
    return a.astype(np.float32) if dtype == torch.float16 else a


SIZES = {
    'cfg5_like': lambda r: r.integers(1, 65, S_MIN + 1000),
    'with_empties': lambda r: r.integers(0, 5, S_MIN + 37),
    # long segments among short ones: whole 2 KB windows inside one segment take the cooperative path
    'long_among_short': lambda r: np.concatenate([r.integers(0, 4, 20000), [5000, 1, 0, 3000, 700], r.integers(1, 9, 14000),
                                                  [9000], r.integers(0, 3, 1000)]),
    'all_ones': lambda r: np.ones(S_MIN + 5, dtype=np.int64),
    'trailing_empties': lambda r: np.concatenate([r.integers(1, 4, S_MIN), np.zeros(70, dtype=np.int64)]),
}


def check(rua, sizes, width, dtype, fn, shift=0):
    n = int(sizes.sum())
    g = torch.Generator().manual_seed(9)
    scale = 0.05 if fn == 'prod' else 1.0
    shape = (n + 1,) if width == 0 else (n + 1, width)
    base = (torch.randn(shape, generator=g) * scale + (1.0 if fn == 'prod' else 0.0)).to(dtype).cuda()
    data = base[shift:shift + n]
    got = to_f32(host(getattr(rua, 'segment_' + fn)(data, torch.from_numpy(sizes).cuda())), dtype)
    x = to_f32(host(data), dtype)
    exp = ora.REDUCERS[fn](x, sizes)
    assert got.shape == exp.shape
    if fn in ('max', 'min'):
        same = (got == exp) | (np.isnan(got) & np.isnan(exp))
        assert same.all(), f'{fn}: {int((~same).sum())} entries differ'
        return
    low = dtype in (torch.bfloat16, torch.float16)
    rtol = 1e-2 if low else (1e-12 if dtype == torch.float64 else 1e-5)
    if fn == 'prod':
        bound = 20 * rtol * np.abs(exp) + 1e-30
    else:
        mag = ora.segment_sum(np.abs(x).astype(np.float64), sizes)
        if fn == 'mean':
            mag = mag / np.maximum(sizes, 1).reshape((-1,) + (1,) * (mag.ndim - 1))
        if fn == 'logsumexp':
            mag = np.ones_like(mag)
        bound = rtol * np.abs(exp) + rtol * mag
    err = np.abs(got.astype(np.float64) - exp.astype(np.float64))
    assert (err <= bound).all(), f'{fn}: worst excess {float((err - bound).max())}'


@pytest.mark.parametrize('case', sorted(SIZES))
@pytest.mark.parametrize('fn', ['sum', 'mean', 'prod', 'max', 'min', 'logsumexp'])
def test_fp32_scalars(rua, case, fn):
    sizes = SIZES[case](np.random.default_rng(5)).astype(np.int64)
    check(rua, sizes, 0, torch.float32, fn)


@pytest.mark.parametrize('dtype,width', [(torch.float32, 2), (torch.float32, 4), (torch.float64, 0), (torch.float64, 2),
                                         (torch.bfloat16, 0), (torch.bfloat16, 2), (torch.bfloat16, 8), (torch.float16, 4)])
@pytest.mark.parametrize('fn', ['sum', 'mean', 'max', 'min', 'logsumexp'])
def test_widths_and_dtypes(rua, dtype, width, fn):
    sizes = SIZES['long_among_short'](np.random.default_rng(6)).astype(np.int64)
    check(rua, sizes, width, dtype, fn)


@pytest.mark.parametrize('fn', ['sum', 'max', 'logsumexp'])
def test_misaligned_base_takes_the_other_kernel_and_agrees(rua, fn):
    sizes = SIZES['cfg5_like'](np.random.default_rng(7)).astype(np.int64)
    check(rua, sizes, 0, torch.float32, fn, shift=1)


def test_quirks_nan_poisoning_and_empty_extreme(rua):
    """reference `initial` semantics (reduce.py:35,40,57-61) through the warp kernel + patch pass."""
    rng = np.random.default_rng(8)
    sizes = rng.integers(0, 4, S_MIN + 11).astype(np.int64)
    n = int(sizes.sum())
    x = torch.randn(n, generator=torch.Generator().manual_seed(1))
    st = torch.from_numpy(sizes).cuda()
    got = rua.segment_max(x.cuda(), st).cpu()
    empty = torch.from_numpy(sizes == 0)
    assert bool((got[empty] == x.min()).all()), 'empty segments of max return the global minimum'
    assert bool((rua.segment_min(x.cuda(), st).cpu()[empty] == x.max()).all())
    assert bool((rua.segment_sum(x.cuda(), st).cpu()[empty] == 0).all())
    assert bool((rua.segment_prod(x.cuda(), st).cpu()[empty] == 1).all())
    assert bool((rua.segment_mean(x.cuda(), st).cpu()[empty] == 0).all())
    x[n // 2] = float('nan')
    for fn in ('max', 'min', 'logsumexp'):
        assert bool(getattr(rua, 'segment_' + fn)(x.cuda(), st).isnan().all()), f'NaN anywhere poisons every output of {fn}'
    s = rua.segment_sum(x.cuda(), st).cpu()
    assert int(s.isnan().sum()) == 1


def test_lengths_that_overrun_the_data_are_clamped(rua):
    """segment_reduce(unsafe=True) semantics are undefined there; the kernel must simply stay inside the array."""
    sizes = np.full(S_MIN, 3, dtype=np.int64)
    x = torch.randn(int(sizes.sum()) - 100, device='cuda')
    out = rua.segment_sum(x, torch.from_numpy(sizes).cuda())
    torch.cuda.synchronize()
    exp = x[:(x.numel() // 3) * 3].view(-1, 3).sum(1)
    assert torch.allclose(out[:exp.numel()], exp, rtol=1e-5, atol=1e-5)


# ---------------------------------------------------------------------------------------------------
# index emit with many short segments (csrc/emit.cu: emit_ptr_warpseg_kernel): C.ptr, L.idx, R.idx, major_sizes_to_ptr
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('case', sorted(SIZES))
def test_ptr_and_idx_many_short_segments(rua, case):
    sizes = SIZES[case](np.random.default_rng(12)).astype(np.int64)
    lens = torch.from_numpy(sizes)
    n, b, t = int(lens.sum()), lens.numel(), int(lens.max())
    which = torch.repeat_interleave(torch.arange(b), lens)
    start = torch.cumsum(lens, 0) - lens
    within = torch.arange(n) - start[which]
    c = rua.C(data=torch.arange(n, device='cuda'), token_sizes=lens.cuda())
    bp, tp = c.ptr()
    assert torch.equal(bp.cpu(), which) and torch.equal(tp.cpu(), within)
    w2, b2 = rua.major_sizes_to_ptr(lens.cuda())
    assert torch.equal(w2.cpu(), within) and torch.equal(b2.cpu(), which)
    if b * t < (1 << 27):
        pad = torch.zeros((b, t), dtype=torch.long, device='cuda')
        left = rua.L(data=pad, token_sizes=lens.cuda())
        assert torch.equal(left.idx().data.cpu(), which * t + within)
        right = rua.R(data=pad, token_sizes=lens.cuda())
        assert torch.equal(right.idx().data.cpu(), which * t + (t - lens[which]) + within)


# ---------------------------------------------------------------------------------------------------
# one-vector rows into C with many short sequences (csrc/rowmap.cu: row_map_warpseg_cat1_kernel): L -> C, R -> C, C.rev, C.roll
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('case', sorted(SIZES))
@pytest.mark.parametrize('dtype,feat', [(torch.long, ()), (torch.int32, ()), (torch.float32, (4,)), (torch.int16, ())])
def test_narrow_rows_into_cat_many_short_sequences(rua, case, dtype, feat):
    sizes = SIZES[case](np.random.default_rng(13)).astype(np.int64)
    lens = torch.from_numpy(sizes).cuda()
    n, b = int(sizes.sum()), sizes.size
    g = torch.Generator(device='cuda').manual_seed(2)
    data = torch.randint(-30000, 30000, (n,) + feat, generator=g, device='cuda').to(dtype)
    which = torch.repeat_interleave(torch.arange(b, device='cuda'), lens)
    start = torch.cumsum(lens, 0) - lens
    within = torch.arange(n, device='cuda') - start[which]
    c = rua.C(data=data, token_sizes=lens)
    assert torch.equal(c.rev().data, data[start[which] + (lens[which] - 1 - within)])
    for shift in (1, -3, 70):
        assert torch.equal(c.roll(shift).data, data[start[which] + (within - shift) % lens[which]]), f'roll({shift})'
    left, right = c.left(7), c.right(7)
    assert torch.equal(left.cat().data, data) and torch.equal(right.cat().data, data)
    assert torch.equal(left.cat().token_sizes, lens)
    # a padded source whose storage is wider than the longest sequence
    wide = rua.L(data=torch.cat([left.data, torch.full_like(left.data[:, :3], -9)], dim=1).contiguous(), token_sizes=lens)
    assert torch.equal(wide.cat().data, data)
