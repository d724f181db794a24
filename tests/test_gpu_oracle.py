"""GPU parity, part 2: the CUDA path against the CPU oracle (oracle/rua_oracle.py, itself pinned to the
reference by tests/test_oracle_golden.py) on seeded random ragged batches -- ragged / empty / single /
tied lengths, odd hidden sizes (every vector width 16/8/4/2/1 bytes), all payload dtypes -- plus
backward passes and size-independent properties at BASELINE.json's full sizes."""
import numpy as np
import pytest
import torch

from oracle import rua_oracle as ora

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def rua():
    import torchrua_b200
    return torchrua_b200


def host(t: torch.Tensor) -> np.ndarray:
    t = t.detach().cpu().contiguous()
    return t.view(torch.uint16).numpy() if t.dtype == torch.bfloat16 else t.numpy()


def to_ora(z):
    if hasattr(z, 'batch_sizes'):
        return ora.Pack(host(z.data), host(z.batch_sizes), host(z.sorted_indices), host(z.unsorted_indices))
    cls = {'CattedSequence': ora.Cat, 'LeftAlignedSequence': ora.Left, 'RightAlignedSequence': ora.Right}[type(z).__name__]
    return cls(host(z.data), host(z.token_sizes))


def same(a: np.ndarray, b: np.ndarray) -> bool:
    return a.shape == b.shape and a.dtype == b.dtype and bool((a == b).all())


def assert_seq_equal(got, exp, what):
    og = to_ora(got)
    assert type(og) is type(exp), what
    assert same(og.data, exp.data), f'{what}: data'
    if isinstance(exp, ora.Pack):
        assert same(og.batch_sizes, exp.batch_sizes), f'{what}: batch_sizes'
        assert same(og.sorted_indices, exp.sorted_indices), f'{what}: sorted_indices'
        assert same(og.unsorted_indices, exp.unsorted_indices), f'{what}: unsorted_indices'
    else:
        assert same(og.token_sizes, exp.token_sizes), f'{what}: token_sizes'


# (lengths, feature shape, dtype)
def length_cases():
    rng = np.random.default_rng(7)
    yield 'single', np.array([1]), (3,), torch.float32
    yield 'one_long', np.array([300]), (5,), torch.float32
    yield 'ties', np.array([4, 4, 4, 2, 2, 7, 7, 1]), (8,), torch.float32
    yield 'zeros_inside', np.array([3, 0, 5, 0, 0, 2, 1]), (4,), torch.float32
    yield 'zeros_tail', np.array([2, 6, 0, 0]), (4,), torch.float32
    yield 'featureless_f32', rng.integers(1, 40, 50), (), torch.float32
    yield 'featureless_i64', rng.integers(1, 40, 33), (), torch.int64
    yield 'featureless_u8', rng.integers(1, 40, 33), (), torch.uint8
    yield 'odd_bytes_1', rng.integers(1, 30, 20), (3,), torch.uint8          # 3-byte rows: 1-byte vectors
    yield 'odd_bytes_2', rng.integers(1, 30, 20), (5,), torch.bfloat16       # 10-byte rows: 2-byte vectors
    yield 'odd_bytes_4', rng.integers(1, 30, 20), (7,), torch.float32        # 28-byte rows: 4-byte vectors
    yield 'odd_bytes_8', rng.integers(1, 30, 20), (6,), torch.float32        # 24-byte rows: 8-byte vectors
    yield 'multi_dim_feat', rng.integers(1, 30, 12), (3, 4), torch.float64
    yield 'wide_bf16', rng.integers(1, 64, 40), (1024,), torch.bfloat16      # 2 KiB rows (bench shape)
    yield 'wide_f16_odd', rng.integers(1, 64, 30), (1000,), torch.float16
    yield 'very_wide', rng.integers(1, 8, 6), (20000,), torch.float32        # column-split path
    yield 'many_short', rng.integers(1, 4, 5000), (2,), torch.int32
    yield 'zipf', np.minimum(rng.zipf(1.5, 300), 500), (16,), torch.float32
    yield 'big_batch', rng.integers(1, 65, 20000), (), torch.int64           # multi-tile scan + 1-pass radix
    yield 'big_batch_2pass', rng.integers(1, 700, 5000), (2,), torch.int16   # 2-pass radix sort


def make(rua, lens: np.ndarray, feat, dtype, seed=0):
    n = int(lens.sum())
    g = torch.Generator().manual_seed(seed)
    if dtype.is_floating_point:
        data = torch.randn((n,) + tuple(feat), generator=g, dtype=torch.float32).to(dtype)
    else:
        hi = 120 if dtype in (torch.uint8, torch.int8) else 30000
        data = torch.randint(1, hi, (n,) + tuple(feat), generator=g).to(dtype)
    c = rua.C(data=data.cuda(), token_sizes=torch.from_numpy(lens.astype(np.int64)).cuda())
    return c, ora.Cat(host(data), lens.astype(np.int64))


def fill_for(dtype, value):
    """the oracle carries bf16 as uint16 bit patterns: translate the fill value accordingly."""
    if dtype == torch.bfloat16:
        return int(ora.f32_to_bf16_bits(np.float32([value]))[0])
    return value


@pytest.mark.parametrize('name,lens,feat,dtype', list(length_cases()), ids=[c[0] for c in length_cases()])
def test_conversions_vs_oracle(rua, name, lens, feat, dtype):
    c, oc = make(rua, lens, feat, dtype)
    ops = {'C': lambda z, f: z.cat(), 'L': lambda z, f: z.left(f), 'R': lambda z, f: z.right(f), 'P': lambda z, f: z.pack()}
    oops = {'C': lambda z, f: ora.to_cat(z), 'L': ora.to_left, 'R': ora.to_right, 'P': lambda z, f: ora.to_pack(z)}
    srcs = {k: ops[k](c, 0) for k in 'CLPR'}
    osrcs = {k: oops[k](oc, 0) for k in 'CLPR'}
    for sk in 'CLPR':
        assert_seq_equal(srcs[sk], osrcs[sk], f'{name}: C->{sk}')
        for dk in 'CLPR':
            for f in ((0, 3) if dk in 'LR' else (0,)):
                got = ops[dk](srcs[sk], f)
                exp = oops[dk](osrcs[sk], fill_for(dtype, f))
                assert_seq_equal(got, exp, f'{name}: {sk}->{dk} fill {f}')


@pytest.mark.parametrize('name,lens,feat,dtype', list(length_cases()), ids=[c[0] for c in length_cases()])
def test_selects_masks_indices_vs_oracle(rua, name, lens, feat, dtype):
    c, oc = make(rua, lens, feat, dtype, seed=1)
    nonempty = bool((lens > 0).all())
    srcs = {'C': c, 'L': c.left(0), 'R': c.right(0), 'P': c.pack()}
    osrcs = {'C': oc, 'L': ora.to_left(oc, 0), 'R': ora.to_right(oc, 0), 'P': ora.to_pack(oc)}
    lo, hi = int(lens.min()), int(lens.max())
    for sk in 'CLPR':
        s, o = srcs[sk], osrcs[sk]
        assert same(host(s.bmask()), ora.bmask(o)), f'{name}: bmask {sk}'
        assert same(host(s.mask(-1, 2, torch.long)), ora.mask(o, -1, 2, np.int64)), f'{name}: mask {sk}'
        assert same(host(rua.get_mask(s)), ora.get_mask(o)), f'{name}: get_mask {sk}'
        b, t = s.ptr()
        ob, ot = ora.ptr(o)
        assert same(host(b), ob) and same(host(t), ot), f'{name}: ptr {sk}'
        assert_seq_equal(s.idx(), ora.idx(o), f'{name}: idx {sk}')
        assert same(host(s.offsets()), ora.offsets(o)), f'{name}: offsets {sk}'
        assert tuple(s.size()) == tuple(ora.size(o)), f'{name}: size {sk}'
        assert_seq_equal(s.rev(), ora.rev(o), f'{name}: rev {sk}')
        for sh in (1, -3, hi + 2):
            assert_seq_equal(s.roll(sh), ora.roll(o, sh), f'{name}: roll({sh}) {sk}')
        if nonempty:
            assert same(host(s.last()), ora.last(o)), f'{name}: last {sk}'
            for n in {1, lo}:
                assert_seq_equal(s.head(n), ora.head(o, n), f'{name}: head({n}) {sk}')
            for a, b_ in {(0, 0), (lo - 1, 0), (0, lo - 1), ((lo - 1) // 2, (lo - 1) - (lo - 1) // 2)}:
                assert_seq_equal(s.trunc((a, b_)), ora.trunc(o, (a, b_)), f'{name}: trunc({a},{b_}) {sk}')


@pytest.mark.parametrize('b,hi', [(1, 5), (2, 3), (31, 9), (33, 9), (1000, 300), (4096, 512), (8192, 70),
                                  (8193, 70), (64, 70000), (300, 5000)])
def test_pack_metadata_paths(rua, b, hi):
    """one-CTA fused metadata kernel (B <= 8192; 1-3 radix passes; T above the speculative cap) and the
    general multi-kernel path (B > 8192) against the oracle's stable descending sort."""
    rng = np.random.default_rng(b * 7 + hi)
    lens = rng.integers(0 if b > 2 else 1, hi + 1, b).astype(np.int64)
    lens[rng.integers(0, b)] = hi          # make sure T == hi
    c = rua.C(data=torch.arange(int(lens.sum()), dtype=torch.int32).cuda(), token_sizes=torch.from_numpy(lens).cuda())
    p = c.pack()
    bs, srt, uns = ora.pack_meta(lens)
    assert same(host(p.batch_sizes), bs) and same(host(p.sorted_indices), srt) and same(host(p.unsorted_indices), uns)
    assert same(host(p.data), ora.to_pack(ora.Cat(np.arange(int(lens.sum()), dtype=np.int32), lens)).data)
    assert same(host(p.cat().data), np.arange(int(lens.sum()), dtype=np.int32))
    assert same(host(c.offsets()), ora.offsets(ora.Cat(host(c.data), lens)))


REDUCE_CASES = [
    # name, sizes, feature shape, dtype
    ('short', lambda r: r.integers(1, 9, 200), (12,), torch.float32),
    ('empties', lambda r: r.integers(0, 4, 300), (8,), torch.float32),
    ('zipf', lambda r: np.minimum(r.zipf(1.5, 400), 4096), (64,), torch.float32),
    ('one_huge', lambda r: np.array([1, 9000, 2]), (32,), torch.float32),
    ('flat', lambda r: r.integers(1, 50, 100), (), torch.float32),
    ('odd_h', lambda r: r.integers(1, 50, 60), (301,), torch.float32),
    ('f64', lambda r: r.integers(1, 300, 40), (6,), torch.float64),
    ('bf16_wide', lambda r: np.minimum(r.zipf(1.5, 200), 2048), (512,), torch.bfloat16),
    ('f16', lambda r: r.integers(1, 100, 50), (40,), torch.float16),
    ('bf16_odd', lambda r: r.integers(1, 100, 50), (7,), torch.bfloat16),
    # average segment below 16 rows: the SHORT instance of the main kernel (csrc/reduce_short.cu), many chunks, rows
    # spanning the whole CTA and rows packed several to a CTA, empty segments, a few long segments crossing many chunks
    ('pooling_bf16', lambda r: r.integers(1, 5, 20000), (256,), torch.bfloat16),
    ('pooling_f32_empties', lambda r: r.integers(0, 4, 30000), (64,), torch.float32),
    ('pooling_f16_wide', lambda r: r.integers(1, 4, 5000), (1024,), torch.float16),
    ('pooling_f64', lambda r: r.integers(1, 6, 8000), (16,), torch.float64),
    ('pieces8_bf16', lambda r: r.integers(7, 10, 4000), (128,), torch.bfloat16),
    ('short_with_giants', lambda r: np.concatenate([r.integers(1, 4, 9000), [5000, 1, 0, 3000], r.integers(0, 3, 9000)]), (96,), torch.float32),
]


def to_f32(a: np.ndarray, dtype) -> np.ndarray:
    if dtype == torch.bfloat16:
        return ora.bf16_bits_to_f32(a)
    if dtype == torch.float16:
        return a.astype(np.float32)
    return a


@pytest.mark.parametrize('name,sizes_fn,feat,dtype', REDUCE_CASES, ids=[c[0] for c in REDUCE_CASES])
@pytest.mark.parametrize('fn', ['sum', 'mean', 'prod', 'max', 'min', 'logsumexp'])
def test_segment_reduce_vs_oracle(rua, name, sizes_fn, feat, dtype, fn):
    """Tolerances (north_star): max/min bit-exact; sum/mean/logsumexp rtol 1e-5 (fp32/fp64) or 1e-2
    (16-bit, against the oracle evaluated on the same values upcast to fp32), each with
    atol = rtol * sum|x| over the segment (sums of randn can cancel to ~0)."""
    sizes = sizes_fn(np.random.default_rng(11)).astype(np.int64)
    n = int(sizes.sum())
    g = torch.Generator().manual_seed(5)
    scale = 0.05 if fn == 'prod' else 1.0
    data = (torch.randn((n,) + feat, generator=g) * scale + (1.0 if fn == 'prod' else 0.0)).to(dtype)
    got = to_f32(host(getattr(rua, 'segment_' + fn)(data.cuda(), torch.from_numpy(sizes).cuda())), dtype)
    x = to_f32(host(data), dtype)
    exp = ora.REDUCERS[fn](x, sizes)
    if fn in ('max', 'min'):
        assert same(got.astype(exp.dtype), exp), f'{name}: {fn}'
        return
    low = dtype in (torch.bfloat16, torch.float16)
    rtol = 1e-2 if low else (1e-12 if dtype == torch.float64 else 1e-5)
    if fn == 'prod':
        rtol *= 20  # products of up to thousands of factors: error grows with length; still relative
        bound = rtol * np.abs(exp) + 1e-30
    else:
        mag = ora.segment_sum(np.abs(x).astype(np.float64), sizes)
        if fn == 'mean':
            mag = mag / np.maximum(sizes, 1).reshape((-1,) + (1,) * (mag.ndim - 1))
        if fn == 'logsumexp':
            mag = np.ones_like(mag)
        bound = rtol * np.abs(exp) + rtol * mag
    err = np.abs(got.astype(np.float64) - exp.astype(np.float64))
    assert (err <= bound).all(), f'{name}: {fn}: worst excess {float((err - bound).max())}'


FLAT_SIZES = {
    'short': lambda r: r.integers(1, 9, 3000),
    'empties': lambda r: r.integers(0, 4, 5000),
    'tile_spanning': lambda r: np.array([5000, 1, 0, 3000, 2, 1025, 1023, 7]),
    'zipf': lambda r: np.minimum(r.zipf(1.5, 2000), 4096),
    'one_row': lambda r: np.array([1]),
    'ragged_tail': lambda r: np.array([3, 1021, 5]),
}


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16, torch.float16, torch.float64])
@pytest.mark.parametrize('case', sorted(FLAT_SIZES))
@pytest.mark.parametrize('fn', ['sum', 'mean', 'prod', 'max', 'min', 'logsumexp'])
def test_flat_segment_reduce_vs_oracle(rua, dtype, case, fn):
    """featureless data (H == 1): the rows-on-lanes kernel with its block-wide segmented scan; also with a
    misaligned base pointer (scalar-load path)."""
    sizes = FLAT_SIZES[case](np.random.default_rng(5)).astype(np.int64)
    n = int(sizes.sum())
    g = torch.Generator().manual_seed(9)
    scale = 0.05 if fn == 'prod' else 1.0
    base = (torch.randn((n + 1,), generator=g) * scale + (1.0 if fn == 'prod' else 0.0)).to(dtype).cuda()
    low = dtype in (torch.bfloat16, torch.float16)
    for shift in (0, 1):
        data = base[shift:shift + n]
        got = to_f32(host(getattr(rua, 'segment_' + fn)(data, torch.from_numpy(sizes).cuda())), dtype)
        x = to_f32(host(data), dtype)
        exp = ora.REDUCERS[fn](x, sizes)
        assert got.shape == exp.shape
        if fn in ('max', 'min'):
            assert same(got.astype(exp.dtype), exp), f'{case}: {fn} shift {shift}'
            continue
        rtol = 1e-2 if low else (1e-12 if dtype == torch.float64 else 1e-5)
        if fn == 'prod':
            bound = 20 * rtol * np.abs(exp) + 1e-30
        else:
            mag = ora.segment_sum(np.abs(x).astype(np.float64), sizes)
            if fn == 'mean':
                mag = mag / np.maximum(sizes, 1)
            if fn == 'logsumexp':
                mag = np.ones_like(mag)
            bound = rtol * np.abs(exp) + rtol * mag
        err = np.abs(got.astype(np.float64) - exp.astype(np.float64))
        assert (err <= bound).all(), f'{case}: {fn} shift {shift}: worst excess {float((err - bound).max())}'


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16, torch.float16, torch.float64])
@pytest.mark.parametrize('width', [2, 4, 8, 3])
@pytest.mark.parametrize('fn', ['sum', 'mean', 'max', 'min', 'logsumexp'])
def test_narrow_segment_reduce_vs_oracle(rua, dtype, width, fn):
    """rows of 2/4/8 elements (<= 16 bytes: rows-on-lanes kernels) and 3 (generic path) against the oracle."""
    rng = np.random.default_rng(21)
    sizes = np.concatenate([rng.integers(0, 6, 700), [3000, 1, 0, 1500], rng.integers(1, 40, 100)]).astype(np.int64)
    n = int(sizes.sum())
    data = torch.randn((n, width), generator=torch.Generator().manual_seed(13)).to(dtype).cuda()
    got = to_f32(host(getattr(rua, 'segment_' + fn)(data, torch.from_numpy(sizes).cuda())), dtype)
    x = to_f32(host(data), dtype)
    exp = ora.REDUCERS[fn](x, sizes)
    if fn in ('max', 'min'):
        assert same(got.astype(exp.dtype), exp)
        return
    low = dtype in (torch.bfloat16, torch.float16)
    rtol = 1e-2 if low else (1e-12 if dtype == torch.float64 else 1e-5)
    mag = ora.segment_sum(np.abs(x).astype(np.float64), sizes)
    if fn == 'mean':
        mag = mag / np.maximum(sizes, 1)[:, None]
    if fn == 'logsumexp':
        mag = np.ones_like(mag)
    err = np.abs(got.astype(np.float64) - exp.astype(np.float64))
    assert (err <= rtol * np.abs(exp) + rtol * mag).all()


def test_segment_head_last_vs_oracle(rua):
    rng = np.random.default_rng(3)
    sizes = rng.integers(1, 30, 100).astype(np.int64)
    data = torch.randn((int(sizes.sum()), 9), generator=torch.Generator().manual_seed(0))
    d, s = data.cuda(), torch.from_numpy(sizes).cuda()
    assert same(host(rua.segment_head(d, s)), ora.segment_head(host(data), sizes))
    assert same(host(rua.segment_last(d, s)), ora.segment_last(host(data), sizes))


# ---------------------------------------------------------------------------------------------------
# backward passes
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('sk', 'CLPR')
@pytest.mark.parametrize('dk', 'CLPR')
def test_conversion_backward(rua, sk, dk):
    """grad of a conversion = the inverse row map applied to the cotangent, zero on padding."""
    lens = np.array([3, 1, 4, 1, 5, 2], dtype=np.int64)
    x = torch.randn((int(lens.sum()), 5), generator=torch.Generator().manual_seed(2))
    leaf = x.cuda().requires_grad_(True)
    c = rua.C(data=leaf, token_sizes=torch.from_numpy(lens).cuda())
    build = {'C': lambda z: z.cat(), 'L': lambda z: z.left(0), 'R': lambda z: z.right(0), 'P': lambda z: z.pack()}
    out = build[dk](build[sk](c))
    w = torch.randn(out.data.shape, generator=torch.Generator().manual_seed(3)).cuda()
    (out.data * w).sum().backward()
    # oracle: route w back through the same conversions
    oc = ora.Cat(host(x), lens)
    obuild = {'C': ora.to_cat, 'L': lambda z: ora.to_left(z, 0), 'R': lambda z: ora.to_right(z, 0), 'P': ora.to_pack}
    ow = obuild[dk](obuild[sk](oc))
    ow.data = host(w)
    exp = ora.to_cat(ow).data
    assert same(host(leaf.grad), exp)


@pytest.mark.parametrize('sk', 'CLPR')
def test_select_backward(rua, sk):
    lens = np.array([3, 2, 4, 2, 5], dtype=np.int64)
    x = torch.randn((int(lens.sum()), 3), generator=torch.Generator().manual_seed(4))
    build = {'C': lambda z: z.cat(), 'L': lambda z: z.left(0), 'R': lambda z: z.right(0), 'P': lambda z: z.pack()}
    tl = torch.from_numpy(lens)

    def torch_ref(fn):
        leaf = x.clone().requires_grad_(True)
        pieces = torch.split(leaf, lens.tolist())
        return leaf, torch.cat([fn(p) for p in pieces], dim=0)

    cases = {
        'rev': (lambda z: z.rev().cat().data, lambda p: p.flip(0)),
        'roll': (lambda z: z.roll(2).cat().data, lambda p: p.roll(2, 0)),
        'trunc': (lambda z: z.trunc((1, 0)).cat().data, lambda p: p[1:]),
        'head': (lambda z: z.head(2).cat().data, lambda p: p[:2]),
        'last': (lambda z: z.last(), lambda p: p[-1:]),
    }
    for name, (ours, ref) in cases.items():
        leaf = x.cuda().requires_grad_(True)
        out = ours(build[sk](rua.C(data=leaf, token_sizes=tl.cuda())))
        rleaf, rout = torch_ref(ref)
        assert same(host(out), host(rout)), f'{name} forward {sk}'
        w = torch.randn(rout.shape, generator=torch.Generator().manual_seed(5))
        (out * w.cuda()).sum().backward()
        (rout * w).sum().backward()
        assert same(host(leaf.grad), host(rleaf.grad)), f'{name} backward {sk}'


@pytest.mark.parametrize('fn', ['sum', 'mean', 'prod', 'max', 'min', 'logsumexp'])
def test_segment_reduce_backward(rua, fn):
    """against torch autograd of the per-segment dense formulation (CPU fp32), rtol 1e-5 / atol 1e-6;
    ties in max/min split the gradient evenly like ATen's SegmentReduceBackward0."""
    sizes = np.array([3, 1, 200, 2, 7, 64, 129], dtype=np.int64)
    x = torch.randn((int(sizes.sum()), 6), generator=torch.Generator().manual_seed(6))
    if fn == 'prod':
        x = x * 0.1 + 1.0
    x[0] = x[1]                      # a tie inside segment 0 (max/min gradient is split)
    x[5] = x[6]
    dense = {'sum': lambda p: p.sum(0), 'mean': lambda p: p.mean(0), 'prod': lambda p: p.prod(0),
             'max': lambda p: p.amax(0), 'min': lambda p: p.amin(0), 'logsumexp': lambda p: p.logsumexp(0)}[fn]
    rleaf = x.clone().requires_grad_(True)
    rout = torch.stack([dense(p) for p in torch.split(rleaf, sizes.tolist())])
    w = torch.randn(rout.shape, generator=torch.Generator().manual_seed(7))
    (rout * w).sum().backward()
    leaf = x.cuda().requires_grad_(True)
    out = getattr(rua, 'segment_' + fn)(leaf, torch.from_numpy(sizes).cuda())
    (out * w.cuda()).sum().backward()
    np.testing.assert_allclose(host(out), host(rout), rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(host(leaf.grad), host(rleaf.grad), rtol=2e-5, atol=2e-6)


# ---------------------------------------------------------------------------------------------------
# BASELINE.json full sizes: size-independent properties
# ---------------------------------------------------------------------------------------------------
def test_cfg2_round_trip_properties(rua):
    """configs[1]: B=4096, len~U[1,512], hidden 1024 bf16.  Round trips are the identity, padding holds
    the fill, masks count the lengths, and a digest of a strided sample agrees with the oracle."""
    g = torch.Generator().manual_seed(0)
    lens = torch.randint(1, 513, (4096,), generator=g)
    n = int(lens.sum())
    data = torch.randn((n, 1024), generator=g, dtype=torch.float32).to(torch.bfloat16).cuda()
    c = rua.C(data=data, token_sizes=lens.cuda())
    p = c.pack()
    left = p.left(7)
    right = left.right(-2)
    back = right.cat()
    assert torch.equal(back.data, data) and torch.equal(back.token_sizes, lens.cuda())
    assert torch.equal(p.cat().data, data) and torch.equal(left.pack().data, p.data)
    assert torch.equal(right.left(7).data, left.data)
    m = c.bmask()
    assert torch.equal(m.sum(dim=1), lens.cuda())
    assert bool((left.data[~m] == 7).all()) and bool((right.data[~m.flip(1)] == -2).all())
    assert torch.equal(left.data[m], data)
    assert torch.equal(c.rev().rev().data, data) and torch.equal(p.roll(5).roll(-5).data, p.data)
    # oracle spot check on a sub-batch (first 64 sequences keep the whole pipeline exact and small)
    k = int(lens[:64].sum())
    sub = rua.C(data=data[:k].clone(), token_sizes=lens[:64].cuda())
    osub = ora.Cat(host(data[:k]), lens[:64].numpy())
    assert same(host(sub.pack().data), ora.to_pack(osub).data)
    assert same(host(sub.right(0).data), ora.to_right(osub, 0).data)


def test_cfg3_reduce_properties(rua):
    """configs[2] (scaled to fit a test: B=4096 Zipf(1.5) lengths <= 4096, hidden 4096 bf16):
    linearity sum(2x) = 2 sum(x) exactly, max >= mean, max(-x) = -min(x), and agreement with a torch
    fp32 index_add reference within 1e-2."""
    rng = np.random.default_rng(0)
    sizes = np.minimum(rng.zipf(1.5, 4096), 4096).astype(np.int64)
    n = int(sizes.sum())
    x = torch.randn((n, 4096), generator=torch.Generator().manual_seed(1)).to(torch.bfloat16).cuda()
    s = torch.from_numpy(sizes).cuda()
    ssum, smax, smin = rua.segment_sum(x, s), rua.segment_max(x, s), rua.segment_min(x, s)
    assert torch.equal(rua.segment_sum(x * 2, s).float(), ssum.float() * 2)
    assert torch.equal(rua.segment_max(-x, s), -smin)
    which = torch.repeat_interleave(torch.arange(sizes.size, device='cuda'), s)
    ref = torch.zeros((sizes.size, 4096), device='cuda').index_add_(0, which, x.float())
    mag = torch.zeros((sizes.size, 4096), device='cuda').index_add_(0, which, x.float().abs())
    assert bool(((ssum.float() - ref).abs() <= 1e-2 * ref.abs() + 1e-2 * mag).all())
    ref_max = torch.full((sizes.size, 4096), float('-inf'), device='cuda').index_reduce_(0, which, x.float(), 'amax')
    assert torch.equal(smax.float(), ref_max)
    lse = rua.segment_logsumexp(x, s).float()
    assert bool((lse >= smax.float() - 1e-2).all())
    assert bool((lse <= smax.float() + torch.log(s.float())[:, None] + 7e-2).all())  # half a bf16 ulp at |x|~16


def test_cfg5_index_only_properties(rua):
    """configs[4]: B=1M, len~U[1,64], no feature dim -- offsets, masks, ptr, idx, sort."""
    g = torch.Generator().manual_seed(0)
    lens = torch.randint(1, 65, (1_000_000,), generator=g)
    n = int(lens.sum())
    c = rua.C(data=torch.arange(n, dtype=torch.long).cuda(), token_sizes=lens.cuda())
    off = c.offsets()
    assert torch.equal(off.cpu(), torch.cumsum(lens, 0) - lens)
    b, t = c.ptr()
    assert torch.equal(b.cpu(), torch.repeat_interleave(torch.arange(lens.numel()), lens))
    assert torch.equal((off[b] + t).cpu(), torch.arange(n))
    p = c.pack()
    srt = p.sorted_indices.cpu()
    assert torch.equal(srt, torch.sort(lens, descending=True, stable=True)[1])
    assert torch.equal(p.batch_sizes, (lens[None, :] > torch.arange(64)[:, None]).sum(1))
    assert torch.equal(p.cat().data.cpu(), torch.arange(n))
    m = c.bmask()
    assert m.shape == (1_000_000, 64) and torch.equal(m.sum(1).cpu(), lens)
    left = c.left(-1)
    assert torch.equal(left.idx().data.cpu(), torch.nonzero(m.view(-1).cpu()).view(-1))


@pytest.mark.parametrize('fn', ['max', 'min'])
@pytest.mark.parametrize('feat,dtype', [((8,), torch.float32), ((256,), torch.float32), ((64,), torch.float64), ((5,), torch.float32)])
def test_segment_extreme_backward_ties(rua, fn, feat, dtype):
    """max / min backward with MANY ties (small-integer data): segments inside one chunk are counted by the thread that
    writes their gradient, segments crossing chunk boundaries through the per-chunk counters (csrc/reduce_bwd.cu);
    both against torch autograd of amax / amin per segment (even split among ties, like SegmentReduceBackward0)."""
    rng = np.random.default_rng(17)
    sizes = np.concatenate([rng.integers(0, 5, 300), [700, 1, 0, 333], rng.integers(1, 4, 200), [130]]).astype(np.int64)
    n = int(sizes.sum())
    x = torch.randint(0, 3, (n,) + feat, generator=torch.Generator().manual_seed(8)).to(dtype)
    dense = (lambda p: p.amax(0)) if fn == 'max' else (lambda p: p.amin(0))
    rleaf = x.clone().requires_grad_(True)
    pieces = [dense(p) if p.shape[0] else torch.zeros(feat, dtype=dtype) for p in torch.split(rleaf, sizes.tolist())]
    rout = torch.stack(pieces)
    w = torch.randn(rout.shape, generator=torch.Generator().manual_seed(9)).to(dtype)
    w[torch.from_numpy(sizes == 0)] = 0            # empty segments: the reference's `initial` value, no gradient
    (rout * w).sum().backward()
    leaf = x.cuda().requires_grad_(True)
    out = getattr(rua, 'segment_' + fn)(leaf, torch.from_numpy(sizes).cuda())
    (out * w.cuda()).sum().backward()
    tol = 1e-12 if dtype == torch.float64 else 1e-6
    np.testing.assert_allclose(host(leaf.grad), host(rleaf.grad), rtol=tol, atol=tol)
