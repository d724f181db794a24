"""Scenarios written against the PUBLIC API only, parameterised by the module (`rua`): the same function runs on the
package under test (in-process) and on the unmodified reference (oracle/ref_worker.py, a subprocess), from the same
seeds, and the two result trees are compared leaf by leaf.  Nothing here imports either package.

Tie order of `pack()` (SURVEY.md 8c hazard 1: the reference sorts non-stably on the CPU): scenarios take
``distinct=True`` to draw pairwise distinct lengths, which makes the permutation unique and every P field
bit-comparable; with ``distinct=False`` a P is reported in canonical form (its `.cat()`, plus `batch_sizes`).
"""
import torch
from torch.nn.utils.rnn import PackedSequence

DT = {'f32': torch.float32, 'f64': torch.float64, 'f16': torch.float16, 'bf16': torch.bfloat16,
      'i64': torch.int64, 'i32': torch.int32, 'u8': torch.uint8, 'bool': torch.bool}
KINDS = 'CLPR'


# ----------------------------------------------------------------------------------------------------------------
# inputs (always drawn on the CPU: identical in both processes) and result flattening
# ----------------------------------------------------------------------------------------------------------------
def make_lengths(seed, B, lo, hi, distinct=False):
    g = torch.Generator().manual_seed(seed)
    if distinct:
        return torch.randperm(B, generator=g) + lo            # lo .. lo+B-1, each once
    return torch.randint(lo, hi + 1, (B,), generator=g)


def make_payload(seed, rows, feat, dtype):
    g = torch.Generator().manual_seed(seed + 1)
    dt = DT[dtype]
    shape = (rows,) + tuple(feat)
    if dt.is_floating_point:
        return torch.randn(shape, generator=g).to(dt)
    return torch.randint(-1000, 1000, shape, generator=g).to(dt)


def make_cat(rua, dev, seed, B, lo, hi, feat, dtype, distinct=False):
    lens = make_lengths(seed, B, lo, hi, distinct)
    data = make_payload(seed, int(lens.sum()), feat, dtype)
    return rua.C(data=data.to(dev), token_sizes=lens.to(dev))


def build(kind, c, fill=0):
    if kind == 'C':
        return c
    if kind == 'P':
        return c.pack()
    return c.left(fill) if kind == 'L' else c.right(fill)


def convert(z, kind, fill=0):
    if kind == 'C':
        return z.cat()
    if kind == 'P':
        return z.pack()
    return z.left(fill) if kind == 'L' else z.right(fill)


def plain(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu()
    if isinstance(x, PackedSequence):
        return ('P', plain(x.data), plain(x.batch_sizes), plain(x.sorted_indices), plain(x.unsorted_indices))
    if hasattr(x, 'token_sizes'):
        return (type(x).__name__, plain(x.data), plain(x.token_sizes))
    if isinstance(x, torch.Size):
        return [int(v) for v in x]
    if isinstance(x, (tuple, list)):
        return [plain(v) for v in x]
    return x


def emit(z, distinct):
    """a sequence result: raw fields when the permutation is unique, canonical form otherwise."""
    if isinstance(z, PackedSequence) and not distinct:
        c = z.cat()
        return ('P~', plain(c.data), plain(c.token_sizes), plain(z.batch_sizes))
    return plain(z)


# ----------------------------------------------------------------------------------------------------------------
# a13-a16: the 12 directed conversions (+ the 4 identities)
# ----------------------------------------------------------------------------------------------------------------
def conversions(rua, dev, seed, B, lo, hi, feat, dtype, distinct=False, fills=(0,)):
    c = make_cat(rua, dev, seed, B, lo, hi, feat, dtype, distinct)
    out = {}
    for sk in KINDS:
        s = build(sk, c)
        for dk in KINDS:
            for f in (fills if dk in 'LR' else (0,)):
                out[f'{sk}->{dk} fill={f}'] = emit(convert(s, dk, f), distinct)
    return out


# ----------------------------------------------------------------------------------------------------------------
# a24-a28: selects
# ----------------------------------------------------------------------------------------------------------------
def selects(rua, dev, seed, B, lo, hi, feat, dtype, distinct=False, shifts=(0, 1, -1, 3), kinds=KINDS):
    c = make_cat(rua, dev, seed, B, lo, hi, feat, dtype, distinct)
    lens = c.token_sizes
    mn, mx = int(lens.min()), int(lens.max())
    out = {}
    for sk in kinds:
        s = build(sk, c)
        out[f'last.{sk}'] = plain(s.last())
        out[f'rev.{sk}'] = emit(s.rev(), distinct)
        for n in sorted({1, mn}):
            out[f'head({n}).{sk}'] = emit(s.head(n), distinct)
        for sh in sorted(set(shifts) | {mx, -mx - 2}):
            out[f'roll({sh}).{sk}'] = emit(s.roll(sh), distinct)
        for a, b in sorted({(0, 0), (mn - 1, 0), (0, mn - 1), ((mn - 1) // 2, (mn - 1) - (mn - 1) // 2)}):
            out[f'trunc({a},{b}).{sk}'] = emit(s.trunc((a, b)), distinct)
    return out


# ----------------------------------------------------------------------------------------------------------------
# a2-a12, a18: index helpers, views, masks.  P entries that depend on the permutation are reported only when it
# is unique (distinct lengths).
# ----------------------------------------------------------------------------------------------------------------
def metadata(rua, dev, seed, B, lo, hi, feat, dtype, distinct=True):
    c = make_cat(rua, dev, seed, B, lo, hi, feat, dtype, distinct)
    out = {}
    sizes = c.token_sizes
    out['get_offsets'] = plain(rua.get_offsets(sizes.clone()))
    out['major_sizes_to_ptr'] = plain(rua.major_sizes_to_ptr(sizes.clone()))
    g = torch.Generator().manual_seed(seed + 3)
    perm = torch.randperm(max(B, 1), generator=g).to(dev)
    out['invert_permutation'] = plain(rua.invert_permutation(perm))
    for sk in KINDS:
        s = build(sk, c)
        perm_free = sk != 'P' or distinct
        out[f'size.{sk}'] = plain(s.size())
        out[f'offsets.{sk}'] = plain(s.offsets())
        out[f'raw.{sk}'] = plain(s.raw()) if perm_free else None
        out[f'get_mask.{sk}'] = plain(rua.get_mask(s))
        out[f'bmask.{sk}'] = plain(s.bmask())
        out[f'mask_long.{sk}'] = plain(s.mask(zero=-1, one=2, dtype=torch.long))
        out[f'mask_f16.{sk}'] = plain(s.mask(zero=-3.5, one=0.25, dtype=torch.float16))
        out[f'mask_default.{sk}'] = plain(s.mask(zero=0, one=1))
        if s.data.is_floating_point():
            out[f'fmask.{sk}'] = plain(s.fmask())
        if perm_free:
            out[f'ptr.{sk}'] = plain(s.ptr())
            out[f'idx.{sk}'] = plain(s.idx())
        # a11 / a12: the four views (incl. the dtype= argument of the padded ones, Appendix B-7)
        out[f'cat_view.{sk}'] = plain(s.cat_view().token_sizes)
        lv = s.left_view(fill_value=3)
        out[f'left_view.{sk}'] = [plain(lv.token_sizes)] + ([] if sk == 'L' else [plain(lv.data)])
        rv = s.right_view(fill_value=-2, dtype=torch.float16)
        out[f'right_view_f16.{sk}'] = [plain(rv.token_sizes)] + ([] if sk == 'R' else [plain(rv.data)])
        lv = s.left_view(fill_value=1, dtype=torch.long)
        out[f'left_view_long.{sk}'] = [] if sk == 'L' else [str(lv.data.dtype), plain(lv.data)]
        pv = s.pack_view()
        out[f'pack_view.{sk}.batch_sizes'] = plain(pv.batch_sizes)
        out[f'pack_view.{sk}.is_same_storage'] = pv.data.data_ptr() == s.data.data_ptr()
        if distinct:
            out[f'pack_view.{sk}.perm'] = [plain(pv.sorted_indices), plain(pv.unsorted_indices)]
        lens = s.cat_view().token_sizes
        srt = pv.sorted_indices
        out[f'pack_view.{sk}.sorted_lengths'] = plain(lens[srt])          # non-increasing on both sides
        out[f'pack_view.{sk}.perm_inverse_ok'] = bool(
            (pv.unsorted_indices[srt] == torch.arange(srt.numel(), device=srt.device)).all())
    return out


# ----------------------------------------------------------------------------------------------------------------
# a17: __getitem__ / __setitem__ with Z keys, (batch_ptr, token_ptr) keys and flat index tensors
# ----------------------------------------------------------------------------------------------------------------
def _widen(rua, s, extra):
    """an L / R whose storage is wider than its longest sequence (Appendix B-2): T comes from token_sizes."""
    if extra == 0 or not hasattr(s, 'token_sizes') or s.data.dim() < 2 or type(s).__name__.startswith('Catted'):
        return s
    b = s.data.size(0)
    pad = torch.full((b, extra) + tuple(s.data.shape[2:]), 9, dtype=s.data.dtype, device=s.data.device)
    return s._replace(data=torch.cat([s.data, pad], dim=1).contiguous())


def _token_keys(seed, lens, count, unique):
    """`count` valid (batch, token) pairs; all different when `unique` (assignment needs a well defined winner)."""
    g = torch.Generator().manual_seed(seed + 7)
    b_all = torch.repeat_interleave(torch.arange(lens.numel()), lens)
    start = torch.cumsum(lens, 0) - lens
    t_all = torch.arange(int(lens.sum())) - start[b_all]
    n = b_all.numel()
    pick = torch.randperm(n, generator=g)[:min(count, n)] if unique else torch.randint(0, n, (count,), generator=g)
    return b_all[pick], t_all[pick]


def getitem(rua, dev, seed, B, lo, hi, feat, dtype, extra_width=0):
    c = make_cat(rua, dev, seed, B, lo, hi, feat, dtype, distinct=True)
    lens = c.token_sizes.cpu()
    kb, kt = _token_keys(seed, lens, 3 * B + 5, unique=False)
    kb, kt = kb.to(dev), kt.to(dev)
    g = torch.Generator().manual_seed(seed + 11)
    out = {}
    for sk in KINDS:
        s = _widen(rua, build(sk, c, fill=-7), extra_width)
        rows = s.raw().size(0)
        out[f'pair.{sk}'] = plain(s[(kb, kt)])
        out[f'pair2d.{sk}'] = plain(s[(kb[:6].view(2, 3), kt[:6].view(2, 3))])
        out[f'Z_idx.{sk}'] = plain(s[s.idx()])                       # a Z key of the same layout
        key = rua.C(data=torch.randint(0, rows, (int(lens[:3].sum()),), generator=g).to(dev),
                    token_sizes=c.token_sizes[:3])
        out[f'Z_cat.{sk}'] = plain(s[key])                           # a Z key of another layout
        out[f'Z_left.{sk}'] = plain(s[key.left(0)])
        flat = torch.randint(-rows, rows, (2 * B + 1,), generator=g).to(dev)
        out[f'flat.{sk}'] = plain(s[flat])                           # negative indices wrap
        out[f'flat2d.{sk}'] = plain(s[flat[:8].view(4, 2)])
        out[f'field0.{sk}'] = plain(s[0])                            # plain tuple indexing still works
        out[f'tensor[Z].{sk}'] = plain(s.raw()[key])                 # the patched Tensor.__getitem__
    return out


def setitem(rua, dev, seed, B, lo, hi, feat, dtype, extra_width=0):
    c = make_cat(rua, dev, seed, B, lo, hi, feat, dtype, distinct=True)
    lens = c.token_sizes.cpu()
    kb, kt = _token_keys(seed, lens, 2 * B + 1, unique=True)
    k = kb.numel()
    n_key = int(lens[:3].sum())
    value = make_payload(seed + 13, max(k, n_key), feat, dtype).to(dev)
    kb, kt = kb.to(dev), kt.to(dev)
    g = torch.Generator().manual_seed(seed + 17)
    out = {}
    for sk in KINDS:
        base = _widen(rua, build(sk, c, fill=-7), extra_width)
        rows = base.raw().size(0)

        def fresh():
            return base._replace(data=base.data.clone())
        s = fresh()
        s[(kb, kt)] = value[:k]
        out[f'pair=tensor.{sk}'] = plain(s.data)
        s = fresh()
        s[(kb, kt)] = 5
        out[f'pair=scalar.{sk}'] = plain(s.data)
        flat = torch.randperm(rows, generator=g)[:k]
        flat = torch.where(torch.arange(k) % 2 == 0, flat, flat - rows).to(dev)      # every other index negative
        s = fresh()
        s[flat] = value[:flat.numel()]
        out[f'flat=tensor.{sk}'] = plain(s.data)
        key = rua.C(data=torch.randperm(rows, generator=g)[:n_key].to(dev), token_sizes=c.token_sizes[:3])
        s = fresh()
        s[key] = value[:n_key]
        out[f'Z=tensor.{sk}'] = plain(s.data)
        s = fresh()
        s[key.left(0)._replace(data=key.left(0).data)] = -1          # padded key: duplicates of row 0, scalar value
        out[f'Zleft=scalar.{sk}'] = plain(s.data)
        raw = base.raw().clone()
        raw[key] = value[:n_key]                                      # the patched Tensor.__setitem__
        out[f'tensor[Z]=.{sk}'] = plain(raw)
    return out


# ----------------------------------------------------------------------------------------------------------------
# a19-a23: segment reductions, .seg
# ----------------------------------------------------------------------------------------------------------------
OPS = ('sum', 'mean', 'prod', 'max', 'min', 'logsumexp')


def reductions(rua, dev, seed, S, lo, hi, feat, dtype, ops=OPS, upcast=False, grad=False, scale=1.0):
    """upcast=True: evaluate in fp32 on the same 16-bit values (the bf16 contract of SURVEY.md 8c hazard 2)."""
    sizes = make_lengths(seed, S, lo, hi)
    data = make_payload(seed, int(sizes.sum()), feat, dtype) * scale
    if upcast:
        data = data.float()
    data, sizes = data.to(dev), sizes.to(dev)
    out = {'abs_sum': plain(rua.segment_sum(data.abs().double(), sizes))}
    for op in ops:
        x = data.clone().requires_grad_(grad)
        if op == 'prod':
            x = (data * 0.05 + 1.0).clone().requires_grad_(grad)
        y = getattr(rua, 'segment_' + op)(x, sizes)
        out[op] = plain(y)
        if grad:
            w = make_payload(seed + 19, y.size(0), tuple(y.shape[1:]), 'f32').to(dev).to(y.dtype)
            (gx,) = torch.autograd.grad((y * w).sum(), [x])
            out[op + '.grad'] = plain(gx)
    if lo >= 1:
        out['head'] = plain(rua.segment_head(data, sizes))
        out['last'] = plain(rua.segment_last(data, sizes))
    return out


def seg(rua, dev, seed, B, lo, hi, feat, dtype, fns=('sum', 'mean', 'max', 'min', 'logsumexp', 'last')):
    c = make_cat(rua, dev, seed, B, lo, hi, feat, dtype, distinct=True)
    g = torch.Generator().manual_seed(seed + 23)
    durations = []
    for n in c.token_sizes.cpu().tolist():
        cuts = torch.unique(torch.randint(n, (n,), generator=g), sorted=False, return_counts=True)[1]
        durations.append(cuts.to(dev))
    d = rua.C.new(durations)
    out = {}
    for sk in KINDS:
        s = build(sk, c)
        for dk in KINDS:
            dd = build(dk, d)
            for fn in fns:
                # the result takes the layout of the DURATIONS' owner; a P result is ordered by the number of
                # durations per sequence, which ties -> canonical form
                out[f'{sk}.seg({dk}, {fn})'] = emit(s.seg(dd, getattr(rua, 'segment_' + fn)), False)
    return out


# ----------------------------------------------------------------------------------------------------------------
# gradients of conversions and selects (IndexBackward / IndexPutBackward in the reference)
# ----------------------------------------------------------------------------------------------------------------
def gradients(rua, dev, seed, B, lo, hi, feat):
    lens = make_lengths(seed, B, lo, hi, distinct=True)
    base = make_payload(seed, int(lens.sum()), feat, 'f32').to(dev)
    lens_d = lens.to(dev)
    mn = int(lens.min())
    out = {}

    def run(tag, fn):
        x = base.clone().requires_grad_(True)
        y = fn(rua.C(data=x, token_sizes=lens_d))
        y = y.data if hasattr(y, 'data') and not isinstance(y, torch.Tensor) else y
        w = make_payload(seed + 29, y.size(0), tuple(y.shape[1:]), 'f32').to(dev)
        (gx,) = torch.autograd.grad((y * w).sum(), [x])
        out[tag] = plain(gx)

    for sk in KINDS:
        for dk in KINDS:
            run(f'{sk}->{dk}', lambda c, sk=sk, dk=dk: convert(build(sk, c), dk, 1.5))
        run(f'last.{sk}', lambda c, sk=sk: build(sk, c).last())
        run(f'rev.{sk}', lambda c, sk=sk: build(sk, c).rev())
        run(f'roll.{sk}', lambda c, sk=sk: build(sk, c).roll(2))
        run(f'head.{sk}', lambda c, sk=sk: build(sk, c).head(mn))
        run(f'trunc.{sk}', lambda c, sk=sk: build(sk, c).trunc((mn // 2, (mn - 1) - mn // 2)))
        run(f'getitem.{sk}', lambda c, sk=sk: build(sk, c)[build(sk, c).idx().rev()])
    return out


# ----------------------------------------------------------------------------------------------------------------
# (f) rows: constructors, compose, scatter
# ----------------------------------------------------------------------------------------------------------------
def constructors(rua, dev, seed, B, lo, hi, feat, dtype):
    lens = make_lengths(seed, B, lo, hi, distinct=True)
    tensors = [make_payload(seed + i, int(n), feat, dtype).to(dev) for i, n in enumerate(lens.tolist())]
    out = {'C.new': plain(rua.C.new(tensors)), 'L.new': plain(rua.L.new(tensors, 2)),
           'R.new': plain(rua.R.new(tensors, -1)), 'P.new': plain(rua.P.new(tensors))}
    for sk in KINDS:
        s = getattr(rua, sk).new(tensors)
        out[f'split.{sk}'] = plain(list(s.split()))
    return out


def compose(rua, dev, seed, B, lo, hi, feat, dtype):
    batches = []
    for k, kind in enumerate(KINDS):
        c = make_cat(rua, dev, seed + 100 * k, B + k, lo, hi, feat, dtype, distinct=True)
        batches.append(build(kind, c))
    p = rua.compose(batches)
    c = p.cat()
    return {'batch_sizes': plain(p.batch_sizes), 'cat.data': plain(c.data), 'cat.token_sizes': plain(c.token_sizes),
            'perm_inverse_ok': bool((p.unsorted_indices[p.sorted_indices] ==
                                     torch.arange(p.sorted_indices.numel(), device=dev)).all())}


def scatter(rua, dev, seed, M, K, feat, dtype):
    g = torch.Generator().manual_seed(seed)
    index = torch.randint(0, max(M - 2, 1), (K,), generator=g).to(dev)
    tensor = make_payload(seed, M, feat, dtype).to(dev)
    source = make_payload(seed + 5, K, feat, dtype).to(dev)
    out = {'abs_sum': plain(rua.scatter_sum(tensor.abs().double(), index, source.abs().double(), include_self=True))}
    for op in OPS:
        for inc in (False, True):
            src = source * 0.05 + 1.0 if op == 'prod' else source
            out[f'{op}.{int(inc)}'] = plain(getattr(rua, 'scatter_' + op)(tensor, index, src, include_self=inc))
    return out


# ----------------------------------------------------------------------------------------------------------------
# timing scenario used by bench.py's reference_cuda record: the bench step through the public API
# ----------------------------------------------------------------------------------------------------------------
def bench_step_outputs(rua, dev, seed, B, lo, hi, feat, dtype):
    """C->P->L->R->C + segment_sum + segment_max on a scaled-down configs[1] batch; results for comparison."""
    c = make_cat(rua, dev, seed, B, lo, hi, feat, dtype)
    back = c.pack().left(0).right(0).cat()
    return {'back': plain(back), 'sum': plain(rua.segment_sum(back.data.float(), back.token_sizes)),
            'max': plain(rua.segment_max(back.data, back.token_sizes))}
