"""Pin the CPU oracle (oracle/rua_oracle.py) to the golden vectors the live reference produced
(tests/golden/make_golden.py).  CPU only; no CUDA involved."""
import numpy as np
import pytest

from oracle import rua_oracle as ora
from tests.helpers import Golden

LAYOUT_CASES = ['small_f32', 'featureless_i64', 'cfg1_f32', 'small_bf16']


def sources(g: Golden):
    """the four layouts of the case's sequence, built by the oracle with the reference's permutation."""
    c = g.seq('src.C', 'C')
    pi = g['src.P.sorted_indices']
    return {
        'C': c,
        'L': ora.to_left(c, 0),
        'R': ora.to_right(c, 0),
        'P': ora.to_pack(c, sorted_indices=pi),
    }, pi


def check_seq(g: Golden, prefix: str, z):
    g.check(prefix + '.data', z.data)
    if isinstance(z, ora.Pack):
        g.check(prefix + '.batch_sizes', z.batch_sizes)
        g.check(prefix + '.sorted_indices', z.sorted_indices)
        g.check(prefix + '.unsorted_indices', z.unsorted_indices)
    else:
        g.check(prefix + '.token_sizes', z.token_sizes)


@pytest.mark.parametrize('case', LAYOUT_CASES)
def test_conversions(case):
    g = Golden(case)
    srcs, pi = sources(g)
    for sk, s in srcs.items():
        check_seq(g, f'src.{sk}', s)
        for dk in 'CLPR':
            if dk in 'LR':
                fills = sorted({int(k.split('fill')[1].split('.')[0]) for k in g.names(f'conv.{sk}{dk}.fill')})
                for f in fills:
                    # bf16 payloads are carried as uint16 bit patterns: the fill is a bit pattern too
                    fv = int(ora.f32_to_bf16_bits(np.float32([f]))[0]) if case == 'small_bf16' else f
                    check_seq(g, f'conv.{sk}{dk}.fill{f}', ora.convert(s, dk, fill_value=fv))
            else:
                check_seq(g, f'conv.{sk}{dk}', ora.convert(s, dk, sorted_indices=pi))


@pytest.mark.parametrize('case', LAYOUT_CASES)
def test_index_helpers_and_masks(case):
    g = Golden(case)
    srcs, _ = sources(g)
    for sk, s in srcs.items():
        g.check(f'size.{sk}', np.asarray(ora.size(s), dtype=np.int64))
        g.check(f'offsets.{sk}', ora.offsets(s))
        b, t = ora.ptr(s)
        g.check(f'ptr.{sk}.batch', b)
        g.check(f'ptr.{sk}.token', t)
        check_seq(g, f'idx.{sk}', ora.idx(s))
        g.check(f'get_mask.{sk}', ora.get_mask(s))
        g.check(f'bmask.{sk}', ora.bmask(s))
        if g.has(f'fmask.{sk}'):
            dt = np.float32 if case != 'small_bf16' else None
            if dt is not None:
                g.check(f'fmask.{sk}', ora.fmask(s, dt))
        g.check(f'mask_long.{sk}', ora.mask(s, -1, 2, np.int64))
        g.check(f'mask_f16.{sk}', ora.mask(s, np.finfo(np.float16).min, np.finfo(np.float16).max, np.float16))
        g.check(f'mask_f64.{sk}', ora.mask(s, np.finfo(np.float64).min, np.finfo(np.float64).max, np.float64))


def test_fmask_bf16_bits():
    g = Golden('small_bf16')
    srcs, _ = sources(g)
    for sk, s in srcs.items():
        m = ora.mask(s, 0xFF7F, 0, np.uint16)  # finfo(bfloat16).min == 0xFF7F
        g.check(f'fmask.{sk}', m)


@pytest.mark.parametrize('case', LAYOUT_CASES)
def test_selects(case):
    g = Golden(case)
    srcs, _ = sources(g)
    for sk, s in srcs.items():
        g.check(f'last.{sk}', ora.last(s))
        check_seq(g, f'rev.{sk}', ora.rev(s))
        for key in g.names('head'):
            if key.endswith(f'.{sk}.data'):
                n = int(key[len('head'):].split('.')[0])
                check_seq(g, f'head{n}.{sk}', ora.head(s, n))
        for key in g.names('roll'):
            if key.endswith(f'.{sk}.data'):
                sh = int(key[len('roll'):].split('.')[0])
                check_seq(g, f'roll{sh}.{sk}', ora.roll(s, sh))
        for key in g.names('trunc'):
            if key.endswith(f'.{sk}.data'):
                a, b = key[len('trunc'):].split('.')[0].split('_')
                check_seq(g, f'trunc{a}_{b}.{sk}', ora.trunc(s, (int(a), int(b))))


REDUCE_FNS = ['sum', 'mean', 'prod', 'max', 'min', 'logsumexp']


@pytest.mark.parametrize('tag,case,fns', [
    ('reduce', 'cfg1_f32', REDUCE_FNS + ['head', 'last']),
    ('empty', 'reduce_edge', REDUCE_FNS),
    ('nan', 'reduce_edge', REDUCE_FNS),
    ('flat', 'reduce_edge', REDUCE_FNS + ['head', 'last']),
    ('f64', 'reduce_edge', REDUCE_FNS + ['head', 'last']),
    ('long', 'reduce_edge', REDUCE_FNS),
])
def test_segment_reduce(tag, case, fns):
    g = Golden(case)
    if tag == 'reduce':
        data, sizes = g['src.C.data'], g['src.C.token_sizes']
    else:
        data, sizes = g[f'{tag}.data'], g[f'{tag}.sizes']
    for fn in fns:
        out = ora.REDUCERS[fn](data, sizes)
        if fn == 'logsumexp':
            # exp/log go through libm in numpy vs SLEEF in ATen: last-ulp differences are expected
            g.check(f'{tag}.{fn}', out, exact=False, rtol=2e-6, atol=1e-6)
        else:
            # strict left-to-right accumulation in the storage dtype reproduces ATen bit-for-bit
            g.check(f'{tag}.{fn}', out)


def test_segment_last_empty_segments_wrap():
    g = Golden('reduce_edge')
    g.check('empty.last', ora.segment_last(g['empty.data'], g['empty.sizes']))


def test_segment_reduce_bf16_contract():
    """bf16 inputs: oracle = reference on the same values upcast to fp32, rounded once."""
    g = Golden('reduce_edge')
    bits, sizes = g['bf16.data'], g['bf16.sizes']
    x = ora.bf16_bits_to_f32(bits)
    for fn in REDUCE_FNS:
        out = ora.REDUCERS[fn](x, sizes)
        if fn == 'logsumexp':
            g.check(f'bf16.{fn}.f32', out, exact=False, rtol=2e-6, atol=1e-6)
        else:
            g.check(f'bf16.{fn}.f32', out)
            g.check(f'bf16.{fn}.rounded', ora.f32_to_bf16_bits(out))


def test_offsets_clamp_quirk():
    g = Golden('reduce_edge')
    c = ora.Cat(np.arange(3, dtype=np.float32), np.array([2, 1, 0], dtype=np.int64))
    g.check('clamp.offsets', ora.offsets(c))


def test_get_offsets_empty_raises():
    with pytest.raises(IndexError):
        ora.get_offsets(np.zeros((0,), dtype=np.int64))


def canonical_cat(z):
    return ora.to_cat(z)


def test_seg():
    g = Golden('seg_f32')
    c = g.seq('src.C', 'C')
    d = g.seq('dur.C', 'C')
    kinds = {'C': lambda s: s, 'L': ora.to_left, 'R': ora.to_right, 'P': ora.to_pack}
    for sk in 'CLPR':
        s = kinds[sk](c)
        for dk in 'CLPR':
            dd = kinds[dk](d)
            for fn in ['sum', 'mean', 'max', 'min', 'logsumexp', 'last', 'prod']:
                out = ora.seg(s, dd, ora.REDUCERS[fn])
                prefix = f'seg.{sk}.{dk}.{fn}'
                exact = fn != 'logsumexp'
                if sk == 'P':
                    # the reference re-packs with its own (non-stable) sort: compare canonical forms
                    ref = ora.to_cat(g.seq(prefix, 'P'))
                    got = ora.to_cat(out)
                    assert (ref.token_sizes == got.token_sizes).all()
                    np.testing.assert_allclose(got.data, ref.data, rtol=0 if exact else 2e-6,
                                               atol=0 if exact else 1e-6)
                else:
                    g.check(prefix + '.token_sizes', out.token_sizes)
                    g.check(prefix + '.data', out.data, exact=exact, rtol=2e-6, atol=1e-6)
    for sk in 'CP':
        out = ora.seg(kinds[sk](c), d, ora.REDUCERS['head'])
        ref = ora.to_cat(g.seq(f'seg.{sk}.C.head', sk))
        assert (ora.to_cat(out).data == ref.data).all()


def test_tie_order_contract():
    """SURVEY.md 8c hazard 1: the reference's sorted_indices is not the stable order; the oracle must
    (i) give identical batch_sizes, (ii) agree on the sorted lengths, (iii) be canonically equal, and
    (iv) be bit-identical when handed the reference's permutation."""
    g = Golden('ties_2000')
    c = g.seq('src.C', 'C')
    ref_pi = g['pack.sorted_indices']
    stable = ora.to_pack(c)
    assert (stable.sorted_indices != ref_pi).any(), 'expected the reference sort to be non-stable here'
    g.check('pack.batch_sizes', stable.batch_sizes)                                       # (i)
    assert (c.token_sizes[stable.sorted_indices] == c.token_sizes[ref_pi]).all()          # (ii)
    assert (stable.unsorted_indices[stable.sorted_indices] == np.arange(2000)).all()
    assert (ora.to_cat(stable).data == c.data).all()                                      # (iii)
    injected = ora.to_pack(c, sorted_indices=ref_pi)                                      # (iv)
    g.check('pack.unsorted_indices', injected.unsorted_indices)
    g.check('pack.data', injected.data)
    g.check('pack.roll3.data', ora.roll(injected, 3).data)
    g.check('pack.rev.data', ora.rev(injected).data)
    g.check('pack.last', ora.last(injected))
    g.check('pack.left.data', ora.to_left(injected, -1).data)
