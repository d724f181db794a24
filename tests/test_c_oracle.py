"""Pin the C restatement (oracle/rua_oracle.c) to the numpy oracle and to the reference's golden vectors.
CPU only."""
import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import rua_oracle as ora
from tests.helpers import Golden


def layouts(c: ora.Cat, pi=None):
    p = ora.to_pack(c, sorted_indices=pi)
    return {'C': c, 'L': ora.to_left(c, 0), 'R': ora.to_right(c, 0), 'P': p}, p


@pytest.mark.parametrize('case', ['small_f32', 'featureless_i64', 'small_bf16', 'cfg1_f32'])
def test_c_oracle_conversions_and_selects(case):
    g = Golden(case)
    c = g.seq('src.C', 'C')
    pi = g['src.P.sorted_indices']
    bs, srt, uns = co.pack_meta(c.token_sizes, pi)
    srcs, p = layouts(c, pi)
    assert (bs == p.batch_sizes).all() and (srt == p.sorted_indices).all() and (uns == p.unsorted_indices).all()
    assert (co.lengths_from_pack(bs, uns) == c.token_sizes).all()
    sbs, ssrt, _ = co.pack_meta(c.token_sizes)
    assert (ssrt == np.argsort(-c.token_sizes, kind='stable')).all() and (sbs == bs).all()
    for sk, s in srcs.items():
        for dk in 'CLPR':
            got = co.move(s.data, sk, dk, c.token_sizes, bs, uns, fill=0)
            assert (got == srcs[dk].data).all(), f'{case}: {sk}->{dk}'
        g.check(f'rev.{sk}.data', co.move(s.data, sk, sk, c.token_sizes, bs, uns, mapping='rev'))
        for sh in (1, -1, 3):
            g.check(f'roll{sh}.{sk}.data', co.move(s.data, sk, sk, c.token_sizes, bs, uns, mapping='roll', shift=sh,
                                                   pad_row0=sk in 'LR'))
    fill = 7 if case != 'small_bf16' else int(ora.f32_to_bf16_bits(np.float32([2]))[0])
    key = 'conv.CL.fill7.data' if case != 'small_bf16' else 'conv.CL.fill2.data'
    if g.has(key):
        g.check(key, co.move(c.data, 'C', 'L', c.token_sizes, fill=fill))
    g.check('bmask.C', co.mask(c.token_sizes, int(c.token_sizes.max()), False, True, np.bool_))


@pytest.mark.parametrize('tag,case', [('reduce', 'cfg1_f32'), ('empty', 'reduce_edge'), ('nan', 'reduce_edge'),
                                      ('long', 'reduce_edge')])
def test_c_oracle_segment_reduce(tag, case):
    g = Golden(case)
    data, sizes = (g['src.C.data'], g['src.C.token_sizes']) if tag == 'reduce' else (g[f'{tag}.data'], g[f'{tag}.sizes'])
    for fn in ['sum', 'mean', 'prod', 'max', 'min', 'logsumexp']:
        got = co.segment_reduce(data, sizes, fn)
        if fn == 'logsumexp':
            g.check(f'{tag}.{fn}', got, exact=False, rtol=2e-6, atol=1e-6)
        else:
            g.check(f'{tag}.{fn}', got)


def test_c_oracle_bf16_contract():
    g = Golden('reduce_edge')
    bits, sizes = g['bf16.data'], g['bf16.sizes']
    for fn in ['sum', 'mean', 'max', 'min', 'prod']:
        g.check(f'bf16.{fn}.rounded', co.segment_reduce(bits, sizes, fn, bf16=True))
