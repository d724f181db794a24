"""SURVEY.md 8(f)-4 fused consumer patterns and 8(f)-2 compose in one launch.

left_mask / left_bmask / left_fmask: padded data + mask from ONE decode (rua_row_map_mask) must be bit-identical to the
separate `X.left(fill)` and `X.mask(...)` calls (which the golden / live-reference tests pin to the reference).
compose: the multi-source gather (rua_gather_rows_multi) must give what `torch.cat(data)[indices]` gives
(torchrua/compose.py:33), forward and backward."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def rua():
    import torchrua_b200
    return torchrua_b200


def build(rua, kind, c):
    return {'C': lambda: c, 'L': lambda: c.left(0), 'R': lambda: c.right(0), 'P': lambda: c.pack()}[kind]()


@pytest.mark.parametrize('kind', 'CLPR')
@pytest.mark.parametrize('feat,dtype', [((64,), torch.float32), ((1024,), torch.bfloat16), ((3, 24), torch.float16),
                                        ((8,), torch.float32), ((), torch.float32), ((33,), torch.float64)])
def test_left_mask_equals_left_plus_mask(rua, kind, feat, dtype):
    g = torch.Generator().manual_seed(3)
    lens = torch.randint(1, 40, (57,), generator=g)
    data = torch.randn((int(lens.sum()),) + feat, generator=g).to(dtype).cuda()
    z = build(rua, kind, rua.C(data=data, token_sizes=lens.cuda()))
    for fill, zero, one, mdt in ((0, False, True, torch.bool), (-2.5, -1, 3, torch.long), (1, 0.5, -0.25, torch.float16)):
        left, m = z.left_mask(fill, zero=zero, one=one, dtype=mdt)
        ref_left, ref_m = z.left(fill), z.mask(zero=zero, one=one, dtype=mdt)
        assert type(left) is type(ref_left) and torch.equal(left.token_sizes, ref_left.token_sizes)
        assert left.data.shape == ref_left.data.shape and torch.equal(left.data, ref_left.data)
        assert m.dtype == ref_m.dtype and m.shape == ref_m.shape and torch.equal(m, ref_m)
    left, m = z.left_bmask(7)
    assert torch.equal(left.data, z.left(7).data) and torch.equal(m, z.bmask())
    left, m = z.left_fmask()
    assert torch.equal(left.data, z.left(0).data) and torch.equal(m, z.fmask())
    cu = z.cu_seqlens()
    assert cu.dtype == torch.int32 and torch.equal(cu.cpu().long(), torch.cat([torch.zeros(1, dtype=torch.long), lens.cumsum(0)]))


def test_left_mask_keeps_gradients(rua):
    lens = torch.tensor([3, 1, 4]).cuda()
    x = torch.randn(8, 64, device='cuda', requires_grad=True)
    left, m = rua.C(data=x, token_sizes=lens).left_fmask()
    w = torch.randn_like(left.data)
    (left.data * w).sum().backward()
    bp, tp = rua.C(data=x.detach(), token_sizes=lens).ptr()
    assert torch.equal(x.grad, w[bp, tp])


def test_left_mask_cfg2_shape(rua):
    """BASELINE configs[1] shape: the fused launch writes B*T mask bytes on top of the conversion's traffic."""
    g = torch.Generator().manual_seed(0)
    lens = torch.randint(1, 513, (4096,), generator=g)
    data = torch.randn((int(lens.sum()), 1024), device='cuda').to(torch.bfloat16)
    c = rua.C(data=data, token_sizes=lens.cuda())
    for z in (c, c.pack()):
        left, m = z.left_fmask()
        assert torch.equal(left.data, z.left(0).data) and torch.equal(m, z.fmask())


@pytest.mark.parametrize('feat,dtype', [((16,), torch.float32), ((256,), torch.bfloat16), ((), torch.long)])
def test_compose_is_cat_then_index(rua, feat, dtype):
    """compose(batches) == pack of all sequences of all batches; the payload rows come straight from every source
    (C, L, P, R storage) in one launch."""
    from torch.nn.utils.rnn import pack_sequence
    g = torch.Generator().manual_seed(5)
    batches, seqs = [], []
    for k, kind in enumerate('CLPRC'):
        lens = torch.randint(1, 9, (4 + k,), generator=g)
        n = int(lens.sum())
        data = (torch.randn((n,) + feat, generator=g) * 10).to(dtype).cuda()
        seqs += list(data.split(lens.tolist()))
        batches.append(build(rua, kind, rua.C(data=data, token_sizes=lens.cuda())))
    p = rua.compose(batches)
    # canonical form: p.cat() lists the sequences in the order of the OUTER packing (compose.py:22-26: first sequence of
    # every batch, batches by descending size, then the second ones, ...)
    counts = [len(z.cat().token_sizes) for z in batches]
    starts = [sum(counts[:k]) for k in range(len(counts))]
    outer = pack_sequence([torch.arange(a, a + n) for a, n in zip(starts, counts)], enforce_sorted=False).data.tolist()
    c = p.cat()
    assert torch.equal(c.data, torch.cat([seqs[j] for j in outer], dim=0))
    assert torch.equal(c.token_sizes.cpu(), torch.tensor([seqs[j].shape[0] for j in outer]))
    ref = pack_sequence(seqs, enforce_sorted=False)
    assert torch.equal(p.batch_sizes, ref.batch_sizes)


def test_compose_backward_reaches_every_source(rua):
    g = torch.Generator().manual_seed(6)
    leaves, batches = [], []
    for k, kind in enumerate('CLPR'):
        lens = torch.randint(1, 6, (3 + k,), generator=g)
        x = torch.randn((int(lens.sum()), 32), generator=g).cuda().requires_grad_(True)
        leaves.append(x)
        batches.append(build(rua, kind, rua.C(data=x, token_sizes=lens.cuda())))
    p = rua.compose(batches)
    w = torch.randn_like(p.data)
    (p.data * w).sum().backward()
    # every token appears exactly once in the result: its gradient is the weight at its packed position
    # the same weights pushed through the INVERSE route: w as a P -> cat (outer order) -> per-sequence pieces
    from torch.nn.utils.rnn import pack_sequence
    wp = rua.P(data=w, batch_sizes=p.batch_sizes, sorted_indices=p.sorted_indices, unsorted_indices=p.unsorted_indices)
    pieces = wp.split()
    counts = [len(z.cat().token_sizes) for z in batches]
    starts = [sum(counts[:k]) for k in range(len(counts))]
    outer = pack_sequence([torch.arange(a, a + n) for a, n in zip(starts, counts)], enforce_sorted=False).data.tolist()
    by_seq = {j: piece for j, piece in zip(outer, pieces)}
    for k, x in enumerate(leaves):
        exp = torch.cat([by_seq[j] for j in range(starts[k], starts[k] + counts[k])], dim=0)
        assert torch.equal(x.grad, exp)


def _durations(rua, lens, piece, seed):
    g = torch.Generator().manual_seed(seed)
    cuts = []
    for n in lens.tolist():
        q, left = [], n
        while left > 0:
            k = min(left, int(torch.randint(1, piece + 1, (1,), generator=g)))
            q.append(k)
            left -= k
        cuts.append(torch.tensor(q).cuda())
    return rua.C.new(cuts)


@pytest.mark.parametrize('feat,dtype', [((64,), torch.float32), ((256,), torch.bfloat16), ((), torch.float32), ((5,), torch.float64)])
@pytest.mark.parametrize('dkind', 'CLPR')
def test_pack_seg_gathers_instead_of_unpacking(rua, feat, dtype, dkind):
    """P.seg(duration, segment_*) reduces straight from the packed rows (rua_segment_reduce_gather over P.idx());
    it must give what the reference's composition P -> C, reduce, C -> P (segment.py:32-33) gives, for all six
    reducers and any layout of the durations (max / min / last bit for bit, sums within the reduction tolerances)."""
    g = torch.Generator().manual_seed(11)
    lens = torch.randperm(70, generator=g)[:45] + 1
    data = torch.randn((int(lens.sum()),) + feat, generator=g).to(dtype).cuda()
    c = rua.C(data=data, token_sizes=lens.cuda())
    p = c.pack()
    dur = build(rua, dkind, _durations(rua, lens, 6, 5))
    for fn in ('sum', 'mean', 'prod', 'max', 'min', 'logsumexp', 'last'):
        f = getattr(rua, 'segment_' + fn)
        got = p.seg(dur, f)
        want = c.seg(dur, f).pack()
        assert isinstance(got, rua.P) and torch.equal(got.batch_sizes, want.batch_sizes)
        assert torch.equal(got.sorted_indices, want.sorted_indices) and torch.equal(got.unsorted_indices, want.unsorted_indices)
        assert got.data.dtype == want.data.dtype and got.data.shape == want.data.shape
        if fn in ('max', 'min', 'last'):
            assert torch.equal(got.data, want.data), fn
        else:       # same fp32 / fp64 accumulation, possibly another association order (DESIGN.md section 4 tolerances)
            tol = {torch.float32: 1e-5, torch.float64: 1e-12, torch.bfloat16: 1e-2}[dtype]
            assert torch.allclose(got.data.double(), want.data.double(), rtol=tol, atol=tol), fn
    # a user-defined reducer keeps the generic path
    got = p.seg(dur, lambda t, s: rua.segment_sum(t, s) * 2)
    assert torch.equal(got.data, (c.seg(dur, rua.segment_sum).pack().data * 2))


def test_pack_seg_gradient(rua):
    g = torch.Generator().manual_seed(12)
    lens = torch.randperm(40, generator=g)[:23] + 1
    base = torch.randn((int(lens.sum()), 32), generator=g).cuda()
    dur = _durations(rua, lens, 5, 6)
    for fn in ('sum', 'mean', 'max', 'logsumexp'):
        f = getattr(rua, 'segment_' + fn)
        grads = []
        for fused in (True, False):
            x = base.clone().requires_grad_(True)
            p = rua.C(data=x, token_sizes=lens.cuda()).pack()
            out = p.seg(dur, f) if fused else p.cat().seg(dur, f).pack()
            w = torch.randn(out.data.shape, generator=torch.Generator().manual_seed(13)).cuda()
            (out.data * w).sum().backward()
            grads.append(x.grad)
        assert torch.allclose(grads[0], grads[1], rtol=1e-6, atol=1e-6), fn


@pytest.mark.parametrize('feat,dtype', [((64,), torch.float32), ((256,), torch.bfloat16), ((5,), torch.float64)])
@pytest.mark.parametrize('kind', 'LR')
@pytest.mark.parametrize('dkind', 'CLPR')
def test_padded_seg_gathers_real_rows_only(rua, feat, dtype, kind, dkind):
    """L.seg / R.seg with segment_sum / mean / prod reduce the N real rows through idx() instead of all B x T rows of the
    padded buffer (segment.py:16-28,38-50); result layout, padding values (0 / 0 / 1) and token_sizes must be what the
    reference's composition gives -- the generic path, forced here by wrapping the reducer in a lambda."""
    g = torch.Generator().manual_seed(21)
    lens = torch.randperm(60, generator=g)[:37] + 1
    data = (torch.randn((int(lens.sum()),) + feat, generator=g) * 0.3 + 1).to(dtype).cuda()
    c = rua.C(data=data, token_sizes=lens.cuda())
    z = build(rua, kind, c)
    dur = build(rua, dkind, _durations(rua, lens, 5, 22))
    for fn in ('sum', 'mean', 'prod'):
        f = getattr(rua, 'segment_' + fn)
        got = z.seg(dur, f)
        want = z.seg(dur, lambda t, s, f=f: f(t, s))
        assert type(got) is type(want) and torch.equal(got.token_sizes, want.token_sizes)
        assert got.data.shape == want.data.shape and got.data.dtype == want.data.dtype
        tol = {torch.float32: 1e-5, torch.float64: 1e-12, torch.bfloat16: 1e-2}[dtype]
        assert torch.allclose(got.data.double(), want.data.double(), rtol=tol, atol=tol), fn
    # gradient: padding rows get none, real rows what the composition gives
    if dtype == torch.float32:
        grads = []
        for fused in (True, False):
            x = z.data.detach().clone().requires_grad_(True)
            zz = z._replace(data=x)
            f = rua.segment_mean if fused else (lambda t, s: rua.segment_mean(t, s))
            out = zz.seg(dur, f)
            w = torch.randn(out.data.shape, generator=torch.Generator().manual_seed(23)).cuda()
            (out.data * w).sum().backward()
            grads.append(x.grad)
        assert torch.allclose(grads[0], grads[1], rtol=1e-5, atol=1e-6)
