"""GPU parity at BASELINE.json's FULL sizes against the C oracle (oracle/rua_oracle.c: the closed forms of the reference,
OpenMP, pinned to the reference's golden vectors by tests/test_c_oracle.py):

  configs[1]  B = 4096, len ~ U[1,512], hidden 1024 bf16: all 12 directed conversions, bit-exact;
  configs[2]  B = 16384, Zipf(1.5) lengths <= 4096, hidden 4096 bf16 (13.2 GB): segment sum / mean / max / min /
              logsumexp (max / min bit-exact; the others within the north_star bf16 tolerance 1e-2 of the oracle's
              fp32-accumulate-round-once values) and head(1) / last / roll / rev on C and P, bit-exact.
"""
import numpy as np
import pytest
import torch

from oracle import c_oracle as co

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def rua():
    import torchrua_b200
    co.build()
    co.use_all_cores()
    return torchrua_b200


def bits(t: torch.Tensor) -> np.ndarray:
    return t.detach().contiguous().view(torch.uint16).cpu().numpy()


def same_on_gpu(got: torch.Tensor, exp_bits: np.ndarray) -> bool:
    exp = torch.from_numpy(exp_bits).cuda()
    return got.shape == exp.shape and bool(torch.equal(got.contiguous().view(torch.uint16), exp))


BF16_7, BF16_M2 = 0x40E0, 0xC000      # bit patterns of bf16(7.0) and bf16(-2.0)


def test_cfg2_all_12_conversions_bit_exact_vs_c_oracle(rua):
    g = torch.Generator().manual_seed(0)
    lens = torch.randint(1, 513, (4096,), generator=g)
    n = int(lens.sum())
    data = torch.randn((n, 1024), generator=g, dtype=torch.float32).to(torch.bfloat16)
    ln = lens.numpy()
    bs, srt, uns = co.pack_meta(ln)
    host = {'C': data.view(torch.uint16).numpy()}
    host['P'] = co.move(host['C'], 'C', 'P', ln, bs, uns)
    host['L'] = co.move(host['C'], 'C', 'L', ln, fill=BF16_7)
    host['R'] = co.move(host['C'], 'C', 'R', ln, fill=BF16_7)
    c = rua.C(data=data.cuda(), token_sizes=lens.cuda())
    dev = {'C': c, 'P': c.pack(), 'L': c.left(7), 'R': c.right(7)}
    assert np.array_equal(dev['P'].batch_sizes.numpy(), bs) and np.array_equal(dev['P'].sorted_indices.cpu().numpy(), srt)
    for sk in 'CLPR':
        assert same_on_gpu(dev[sk].data, host[sk]), f'source layout {sk}'
        for dk in 'CLPR':
            if dk == sk:
                continue
            if dk in 'LR':
                got = (dev[sk].left(-2) if dk == 'L' else dev[sk].right(-2)).data
                exp = co.move(host[sk], sk, dk, ln, bs, uns, fill=BF16_M2)
            else:
                got = (dev[sk].cat() if dk == 'C' else dev[sk].pack()).data
                exp = co.move(host[sk], sk, dk, ln, bs, uns)
            assert same_on_gpu(got, exp), f'{sk} -> {dk} differs from the oracle at full size'
            del got, exp


@pytest.fixture(scope='module')
def cfg3(rua):
    rng = np.random.default_rng(0)
    sizes = np.minimum(rng.zipf(1.5, 16384), 4096).astype(np.int64)
    n = int(sizes.sum())
    x = torch.randn((n, 4096), device='cuda', dtype=torch.float32, generator=torch.Generator(device='cuda').manual_seed(1))
    x = x.to(torch.bfloat16)
    return sizes, x, bits(x)


def test_cfg3_reductions_vs_c_oracle(rua, cfg3):
    sizes, x, xb = cfg3
    s = torch.from_numpy(sizes).cuda()
    mag = rua.segment_sum(x.abs(), s).float()                      # sum |x| per (segment, column)
    lens = s.float().clamp_min(1)[:, None]
    for fn in ('sum', 'mean', 'max', 'min', 'logsumexp'):
        got = getattr(rua, 'segment_' + fn)(x, s)
        exp = torch.from_numpy(co.segment_reduce(xb, sizes, fn, bf16=True)).cuda().view(torch.bfloat16)
        assert got.shape == exp.shape and got.dtype == torch.bfloat16
        if fn in ('max', 'min'):
            assert torch.equal(got, exp), f'segment_{fn} is not bit-exact at full size'
            continue
        # north_star: 1e-2 relative for bf16 (+ the same fraction of sum|x| -- a sum of randn can cancel to ~0)
        scale = {'sum': mag, 'mean': mag / lens, 'logsumexp': torch.ones_like(mag)}[fn]
        err = (got.float() - exp.float()).abs()
        bound = 1e-2 * exp.float().abs() + 1e-2 * scale
        assert bool((err <= bound).all()), f'segment_{fn}: {int((err > bound).sum())} entries outside 1e-2'
        # and in practice: at most one bf16 ulp from the oracle's rounding of the same fp32 value
        tight = err <= exp.float().abs() * 2.0 ** -7 + 1e-2 * scale * 2.0 ** -7 + 1e-30
        assert float(tight.float().mean()) > 0.999, f'segment_{fn}: only {float(tight.float().mean()):.4f} within one bf16 ulp'


def test_cfg3_selects_bit_exact_vs_c_oracle(rua, cfg3):
    sizes, x, xb = cfg3
    s = torch.from_numpy(sizes).cuda()
    bs, srt, uns = co.pack_meta(sizes)
    c = rua.C(data=x, token_sizes=s)
    p = c.pack()
    pb = co.move(xb, 'C', 'P', sizes, bs, uns)
    assert same_on_gpu(p.data, pb), 'C -> P at full size'
    # rev / roll keep the layout: the oracle moves C -> C (resp. P -> P) with the token map applied
    assert same_on_gpu(c.rev().data, co.move(xb, 'C', 'C', sizes, mapping='rev')), 'C.rev'
    assert same_on_gpu(p.roll(1).data, co.move(pb, 'P', 'P', sizes, bs, uns, mapping='roll', shift=1)), 'P.roll(1)'
    assert same_on_gpu(c.roll(1).data, co.move(xb, 'C', 'C', sizes, mapping='roll', shift=1)), 'C.roll(1)'
    assert same_on_gpu(p.rev().data, co.move(pb, 'P', 'P', sizes, bs, uns, mapping='rev')), 'P.rev'
    # head(1) / last: one row per sequence
    off = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    first = torch.from_numpy(xb[off]).cuda()
    last = torch.from_numpy(xb[off + sizes - 1]).cuda()
    assert torch.equal(c.head(1).data.view(torch.uint16), first)
    assert torch.equal(c.last().view(torch.uint16), last) and torch.equal(p.last().view(torch.uint16), last)
    ph = p.head(1)
    assert torch.equal(ph.cat().data.view(torch.uint16), first) and ph.batch_sizes.tolist() == [sizes.size]
