"""GPU parity, part 1: the CUDA path against the golden vectors the live reference produced
(tests/golden/*.npz).  Bit-exact for conversions, selects, masks and index helpers; PackedSequence
results are compared with the reference's permutation injected (and in canonical form otherwise).
Everything here goes through the public API, i.e. through the C-ABI in librua_b200.so."""
import numpy as np
import pytest
import torch

from tests.helpers import Golden

pytestmark = pytest.mark.gpu

LAYOUT_CASES = ['small_f32', 'featureless_i64', 'cfg1_f32', 'small_bf16']


@pytest.fixture(scope='module')
def rua():
    import torchrua_b200
    return torchrua_b200


def dev(a: np.ndarray, bf16=False) -> torch.Tensor:
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t.view(torch.bfloat16) if bf16 else t


def host(t: torch.Tensor) -> np.ndarray:
    t = t.detach().cpu().contiguous()
    return t.view(torch.uint16).numpy() if t.dtype == torch.bfloat16 else t.numpy()


def sources(rua, g: Golden, case: str):
    bf16 = case == 'small_bf16'
    c = rua.C(data=dev(g['src.C.data'], bf16), token_sizes=dev(g['src.C.token_sizes']))
    pi = dev(g['src.P.sorted_indices'])
    return {'C': c, 'L': c.left(0), 'R': c.right(0), 'P': c.pack(sorted_indices=pi)}, pi


def check_seq(g: Golden, prefix: str, z):
    g.check(prefix + '.data', host(z.data))
    if hasattr(z, 'batch_sizes'):
        assert not z.batch_sizes.is_cuda, 'batch_sizes must live on the host (PackedSequence contract)'
        g.check(prefix + '.batch_sizes', host(z.batch_sizes))
        g.check(prefix + '.sorted_indices', host(z.sorted_indices))
        g.check(prefix + '.unsorted_indices', host(z.unsorted_indices))
    else:
        g.check(prefix + '.token_sizes', host(z.token_sizes))


@pytest.mark.parametrize('case', LAYOUT_CASES)
def test_conversions(rua, case):
    g = Golden(case)
    srcs, pi = sources(rua, g, case)
    for sk, s in srcs.items():
        check_seq(g, f'src.{sk}', s)
        for dk in 'CLPR':
            if dk in 'LR':
                fills = sorted({int(k.split('fill')[1].split('.')[0]) for k in g.names(f'conv.{sk}{dk}.fill')})
                for f in fills:
                    out = s.left(f) if dk == 'L' else s.right(f)
                    check_seq(g, f'conv.{sk}{dk}.fill{f}', out)
            elif dk == 'C':
                check_seq(g, f'conv.{sk}C', s.cat())
            elif sk == 'P':
                check_seq(g, 'conv.PP', s.pack())
            else:
                check_seq(g, f'conv.{sk}P', s.pack(sorted_indices=pi))


@pytest.mark.parametrize('case', LAYOUT_CASES)
def test_index_helpers_and_masks(rua, case):
    g = Golden(case)
    srcs, _ = sources(rua, g, case)
    for sk, s in srcs.items():
        g.check(f'size.{sk}', np.asarray(s.size(), dtype=np.int64))
        g.check(f'offsets.{sk}', host(s.offsets()))
        b, t = s.ptr()
        g.check(f'ptr.{sk}.batch', host(b))
        g.check(f'ptr.{sk}.token', host(t))
        check_seq(g, f'idx.{sk}', s.idx())
        g.check(f'get_mask.{sk}', host(rua.get_mask(s)))
        g.check(f'bmask.{sk}', host(s.bmask()))
        if g.has(f'fmask.{sk}'):
            g.check(f'fmask.{sk}', host(s.fmask()))
        g.check(f'mask_long.{sk}', host(s.mask(zero=-1, one=2, dtype=torch.long)))
        g.check(f'mask_f16.{sk}', host(s.mask(zero=torch.finfo(torch.float16).min,
                                              one=torch.finfo(torch.float16).max, dtype=torch.float16)))
        g.check(f'mask_f64.{sk}', host(s.mask(zero=torch.finfo(torch.float64).min,
                                              one=torch.finfo(torch.float64).max, dtype=torch.float64)))


@pytest.mark.parametrize('case', LAYOUT_CASES)
def test_selects(rua, case):
    g = Golden(case)
    srcs, _ = sources(rua, g, case)
    for sk, s in srcs.items():
        g.check(f'last.{sk}', host(s.last()))
        check_seq(g, f'rev.{sk}', s.rev())
        for key in g.names('head'):
            if key.endswith(f'.{sk}.data'):
                n = int(key[len('head'):].split('.')[0])
                check_seq(g, f'head{n}.{sk}', s.head(n))
        for key in g.names('roll'):
            if key.endswith(f'.{sk}.data'):
                sh = int(key[len('roll'):].split('.')[0])
                check_seq(g, f'roll{sh}.{sk}', s.roll(sh))
        for key in g.names('trunc'):
            if key.endswith(f'.{sk}.data'):
                a, b = key[len('trunc'):].split('.')[0].split('_')
                check_seq(g, f'trunc{a}_{b}.{sk}', s.trunc((int(a), int(b))))


REDUCE_FNS = ['sum', 'mean', 'prod', 'max', 'min', 'logsumexp']


def seg_abs_sum(data: np.ndarray, sizes: np.ndarray) -> np.ndarray:
    from oracle import rua_oracle as ora
    return ora.segment_sum(np.abs(data).astype(np.float64), sizes)


def check_reduce(g, key, got, fn, data, sizes, rtol=1e-5):
    """north_star tolerance: bit-exact for max/min (order independent); rtol 1e-5 for fp32 sum/mean/
    logsumexp, plus atol = rtol * sum|x| per segment because a randn sum can land arbitrarily close to 0
    (SURVEY.md 8c hazard 2).  prod follows the sum rule on a log scale: plain rtol with a small atol."""
    expected = g[key]
    assert got.shape == expected.shape and got.dtype == expected.dtype, key
    if fn in ('max', 'min', 'head', 'last'):
        same = (got == expected) | (np.isnan(got) & np.isnan(expected))
        assert same.all(), f'{key}: {int((~same).sum())} mismatches'
        return
    if fn == 'prod':
        np.testing.assert_allclose(got, expected, rtol=1e-4, atol=1e-30, equal_nan=True, err_msg=key)
        return
    scale = seg_abs_sum(data, sizes).astype(np.float64)
    if fn == 'mean':
        scale = scale / np.maximum(sizes, 1).reshape((-1,) + (1,) * (scale.ndim - 1))
    if fn == 'logsumexp':
        scale = np.ones_like(scale)
    err = np.abs(got.astype(np.float64) - expected.astype(np.float64))
    bound = rtol * np.abs(expected.astype(np.float64)) + rtol * scale
    ok = (err <= bound) | (np.isnan(got) & np.isnan(expected)) | (got == expected)
    assert ok.all(), f'{key}: max excess {np.nanmax(err - bound)}'


@pytest.mark.parametrize('tag,case,fns', [
    ('reduce', 'cfg1_f32', REDUCE_FNS + ['head', 'last']),
    ('empty', 'reduce_edge', REDUCE_FNS),
    ('nan', 'reduce_edge', REDUCE_FNS),
    ('flat', 'reduce_edge', REDUCE_FNS + ['head', 'last']),
    ('f64', 'reduce_edge', REDUCE_FNS + ['head', 'last']),
    ('long', 'reduce_edge', REDUCE_FNS),
])
def test_segment_reduce(rua, tag, case, fns):
    g = Golden(case)
    if tag == 'reduce':
        data, sizes = g['src.C.data'], g['src.C.token_sizes']
    else:
        data, sizes = g[f'{tag}.data'], g[f'{tag}.sizes']
    d, s = dev(data), dev(sizes)
    for fn in fns:
        got = host(getattr(rua, 'segment_' + fn)(d, s))
        check_reduce(g, f'{tag}.{fn}', got, fn, np.nan_to_num(data), sizes, rtol=1e-5 if data.dtype != np.float64 else 1e-12)


def test_segment_last_empty_segments_wrap(rua):
    g = Golden('reduce_edge')
    g.check('empty.last', host(rua.segment_last(dev(g['empty.data']), dev(g['empty.sizes']))))


def test_segment_reduce_bf16_contract(rua):
    """bf16: fp32 accumulation, rounded once; oracle = reference on the same values upcast to fp32
    (rtol 1e-2 per north_star; in practice max/min are exact and sums differ by at most 1 bf16 ulp)."""
    from oracle import rua_oracle as ora
    g = Golden('reduce_edge')
    bits, sizes = g['bf16.data'], g['bf16.sizes']
    d, s = dev(bits, bf16=True), dev(sizes)
    x = ora.bf16_bits_to_f32(bits)
    for fn in REDUCE_FNS:
        got = ora.bf16_bits_to_f32(host(getattr(rua, 'segment_' + fn)(d, s)))
        exp = g[f'bf16.{fn}.f32']
        if fn in ('max', 'min'):
            assert (ora.f32_to_bf16_bits(got) == g[f'bf16.{fn}.rounded']).all()
            continue
        scale = ora.segment_sum(np.abs(x), sizes)
        if fn == 'mean':
            scale = scale / np.maximum(sizes, 1)[:, None]
        if fn in ('logsumexp', 'prod'):
            scale = np.abs(exp)
        err = np.abs(got - exp)
        assert (err <= 1e-2 * np.abs(exp) + 1e-2 * scale + 1e-30).all(), fn


def test_seg(rua):
    """.seg(duration, fn): 4 sequence layouts x 4 duration layouts x 7 reducers (+ head on C/P)."""
    g = Golden('seg_f32')
    c = rua.C(data=dev(g['src.C.data']), token_sizes=dev(g['src.C.token_sizes']))
    d = rua.C(data=dev(g['dur.C.data']), token_sizes=dev(g['dur.C.token_sizes']))
    build = {'C': lambda z: z, 'L': lambda z: z.left(0), 'R': lambda z: z.right(0), 'P': lambda z: z.pack()}
    for sk in 'CLPR':
        s = build[sk](c)
        for dk in 'CLPR':
            dd = build[dk](d)
            for fn in ['sum', 'mean', 'max', 'min', 'logsumexp', 'last', 'prod']:
                out = s.seg(dd, getattr(rua, 'segment_' + fn))
                prefix = f'seg.{sk}.{dk}.{fn}'
                exact = fn in ('max', 'min', 'last')
                if sk == 'P':
                    got = host(out.cat().data)
                    ref_p = rua.P(data=dev(g[prefix + '.data']), batch_sizes=torch.from_numpy(g[prefix + '.batch_sizes']),
                                  sorted_indices=dev(g[prefix + '.sorted_indices']),
                                  unsorted_indices=dev(g[prefix + '.unsorted_indices']))
                    exp = host(ref_p.cat().data)
                    g.check(prefix + '.batch_sizes', host(out.batch_sizes))
                else:
                    got, exp = host(out.data), g[prefix + '.data']
                    g.check(prefix + '.token_sizes', host(out.token_sizes))
                if exact:
                    assert (got == exp).all(), prefix
                else:
                    np.testing.assert_allclose(got, exp, rtol=1e-5, atol=2e-6, err_msg=prefix)
    for sk in 'CP':
        out = build[sk](c).seg(d, rua.segment_head)
        ref = g[f'seg.C.C.head.data']
        assert (host(out.cat().data) == ref).all()


def test_compose_split_tolist(rua):
    """compose() of four ragged batches in C/L/P/R layouts (-> one PackedSequence), split() and tolist():
    the rows just outside the hot path, compared in canonical (cat) form with the reference's outputs."""
    g = Golden('compose_f32')
    build = {'C': lambda z: z, 'L': lambda z: z.left(0), 'R': lambda z: z.right(0), 'P': lambda z: z.pack()}
    batches = []
    for b, kind in enumerate('CLPR'):
        c = rua.C(data=dev(g[f'in{b}.data']), token_sizes=dev(g[f'in{b}.token_sizes']))
        batches.append(build[kind](c))
    out = rua.compose(batches)
    g.check('out.batch_sizes', host(out.batch_sizes))
    cat = out.cat()
    g.check('out.cat.token_sizes', host(cat.token_sizes))
    g.check('out.cat.data', host(cat.data))
    assert (host(out.unsorted_indices)[host(out.sorted_indices)] == np.arange(out.sorted_indices.numel())).all()
    for k, piece in enumerate(batches[2].split()):
        g.check(f'split2.{k}', host(piece))
    for k, piece in enumerate(batches[1].split()):
        g.check(f'split1.{k}', host(piece))
    g.check('tolist3', np.asarray([len(x) for x in batches[3].tolist()], dtype=np.int64))


@pytest.mark.parametrize('op', ['sum', 'mean', 'prod', 'max', 'min', 'logsumexp'])
@pytest.mark.parametrize('include_self', [False, True])
def test_scatter_reductions(rua, op, include_self):
    """scatter_* (unsorted-index reductions, 'next' row 8f-1): forward and both gradients against the
    reference.  max/min are bit-exact; the sums follow the fp32 tolerance (rtol 1e-5, atol 1e-5)."""
    g = Golden('scatter_f32')
    index = dev(g['index'])
    weight = dev(g['weight'])
    t = dev(g['tensor']).requires_grad_(True)
    src = g['source'] * np.float32(0.3) + np.float32(1.0) if op == 'prod' else g['source']
    s = dev(src).requires_grad_(True)
    out = getattr(rua, 'scatter_' + op)(t, index, s, include_self=include_self)
    tag = f'{op}.{int(include_self)}'
    exact = op in ('max', 'min')
    g.check(tag + '.out', host(out), exact=exact, rtol=1e-5, atol=1e-5)
    gt, gs = torch.autograd.grad((out * weight)[torch.isfinite(out)].sum(), [t, s], allow_unused=True)
    gt = torch.zeros_like(t) if gt is None else gt
    gs = torch.zeros_like(s) if gs is None else gs
    g.check(tag + '.grad_tensor', host(gt), exact=False, rtol=1e-5, atol=1e-5)
    g.check(tag + '.grad_source', host(gs), exact=False, rtol=1e-4, atol=1e-5)
    # other dims and an integer payload (ATen composition, like the reference)
    if op == 'sum':
        out_t = rua.scatter_sum(t.detach().t().contiguous(), index, s.detach().t().contiguous(), include_self, dim=1)
        np.testing.assert_allclose(host(out_t.t()), g[tag + '.out'], rtol=1e-5, atol=1e-5)
        ints = rua.scatter_sum(torch.zeros(9, dtype=torch.long, device='cuda'), index,
                               torch.ones(40, dtype=torch.long, device='cuda'))
        assert int(ints.sum()) == 40


def test_tie_order_contract(rua):
    """SURVEY.md 8c hazard 1 on the CUDA path: (i) batch_sizes bit-equal, (ii) sorted lengths equal and
    the permutation consistent, (iii) canonical equality, (iv) bit-identical data with the reference's
    permutation injected; the default device sort is the STABLE descending order."""
    g = Golden('ties_2000')
    lens = g['src.C.token_sizes']
    c = rua.C(data=dev(g['src.C.data']), token_sizes=dev(lens))
    ref_pi = g['pack.sorted_indices']
    p = c.pack()
    ours = host(p.sorted_indices)
    assert (ours == np.argsort(-lens, kind='stable')).all(), 'device sort must be stable descending'
    assert (ours != ref_pi).any()
    g.check('pack.batch_sizes', host(p.batch_sizes))                              # (i)
    assert (lens[ours] == lens[ref_pi]).all()                                     # (ii)
    assert (host(p.unsorted_indices)[ours] == np.arange(lens.size)).all()
    assert (host(p.cat().data) == g['src.C.data']).all()                          # (iii)
    q = c.pack(sorted_indices=dev(ref_pi))                                        # (iv)
    g.check('pack.unsorted_indices', host(q.unsorted_indices))
    g.check('pack.data', host(q.data))
    g.check('pack.roll3.data', host(q.roll(3).data))
    g.check('pack.rev.data', host(q.rev().data))
    g.check('pack.last', host(q.last()))
    g.check('pack.left.data', host(q.left(-1).data))
