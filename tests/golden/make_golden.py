"""Generate the golden vectors under tests/golden/ from the LIVE reference (speedcell4/torchrua).

Run in the authoring container only (the reference is mounted at /root/reference there and does
not exist on the GPU box):

    cd /tmp && PYTHONDONTWRITEBYTECODE=1 python /root/repo/tests/golden/make_golden.py

Every array written here is an output of the unmodified reference on CPU (torch 2.11.0+cu128).
Large payload outputs are stored as sha256 digests of their raw bytes (``sha:`` prefix); inputs and
small outputs are stored in full.  The oracle (oracle/rua_oracle.py) and the CUDA path are both
checked against these files.
"""
import hashlib
import os
import sys

import numpy as np
import torch

REF = os.environ.get('RUA_REFERENCE', '/root/reference')
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

import torchrua as rua  # noqa: E402  (the reference)
from torchrua import C, L, P, R  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
KINDS = {'C': C, 'L': L, 'P': P, 'R': R}
FULL_LIMIT = 1 << 15  # bytes; larger arrays are stored as digests


def np_of(t: torch.Tensor) -> np.ndarray:
    t = t.detach().cpu().contiguous()
    if t.dtype == torch.bfloat16:
        return t.view(torch.uint16).numpy()
    return t.numpy()


def digest(a: np.ndarray) -> np.ndarray:
    h = hashlib.sha256()
    h.update(str(a.dtype).encode())
    h.update(str(tuple(a.shape)).encode())
    h.update(np.ascontiguousarray(a).tobytes())
    return np.frombuffer(h.digest(), dtype=np.uint8).copy()


class Rec(dict):
    def put(self, name, t, full=False):
        a = np_of(t) if isinstance(t, torch.Tensor) else np.asarray(t)
        if full or a.nbytes <= FULL_LIMIT:
            self[name] = a
        else:
            self['sha:' + name] = digest(a)

    def put_seq(self, name, z, full=False):
        self.put(name + '.data', z.data, full)
        if isinstance(z, P):
            self.put(name + '.batch_sizes', z.batch_sizes, True)
            self.put(name + '.sorted_indices', z.sorted_indices, True)
            self.put(name + '.unsorted_indices', z.unsorted_indices, True)
        else:
            self.put(name + '.token_sizes', z.token_sizes, True)


def build(kind, c, fill=0):
    if kind == 'C':
        return c
    if kind == 'P':
        return c.pack()
    return c.left(fill) if kind == 'L' else c.right(fill)


def layout_case(rec: Rec, c: C, fills=(0, 7, -3)):
    """all 12 directed conversions + metadata helpers + masks + selects."""
    srcs = {k: build(k, c) for k in 'CLPR'}
    for sk, s in srcs.items():
        rec.put_seq(f'src.{sk}', s, full=(sk == 'C'))
        rec.put(f'size.{sk}', np.asarray(s.size(), dtype=np.int64), True)
        rec.put(f'offsets.{sk}', s.offsets(), True)
        b, t = s.ptr()
        rec.put(f'ptr.{sk}.batch', b)
        rec.put(f'ptr.{sk}.token', t)
        rec.put_seq(f'idx.{sk}', s.idx())
        rec.put(f'get_mask.{sk}', rua.get_mask(s))
        rec.put(f'bmask.{sk}', s.bmask())
        if s.data.is_floating_point():
            rec.put(f'fmask.{sk}', s.fmask())
        rec.put(f'mask_long.{sk}', s.mask(zero=-1, one=2, dtype=torch.long))
        rec.put(f'mask_f16.{sk}', s.mask(zero=torch.finfo(torch.float16).min,
                                         one=torch.finfo(torch.float16).max, dtype=torch.float16))
        rec.put(f'mask_f64.{sk}', s.mask(zero=torch.finfo(torch.float64).min,
                                         one=torch.finfo(torch.float64).max, dtype=torch.float64))
        for dk in 'CLPR':
            if dk in 'LR':
                for f in fills:
                    out = getattr(s, {'L': 'left', 'R': 'right'}[dk])(f)
                    rec.put_seq(f'conv.{sk}{dk}.fill{f}', out)
            else:
                out = getattr(s, {'C': 'cat', 'P': 'pack'}[dk])()
                rec.put_seq(f'conv.{sk}{dk}', out)

    lens = c.token_sizes
    lo, hi = int(lens.min()), int(lens.max())
    for sk, s in srcs.items():
        rec.put(f'last.{sk}', s.last())
        rec.put_seq(f'rev.{sk}', s.rev())
        for n in sorted({1, lo}):
            rec.put_seq(f'head{n}.{sk}', s.head(n))
        for sh in sorted({0, 1, -1, 3, -hi, hi, hi + 5, -2 * hi - 1}):
            rec.put_seq(f'roll{sh}.{sk}', s.roll(sh))
        for a, b in sorted({(0, 0), (lo - 1, 0), (0, lo - 1), ((lo - 1) // 2, (lo - 1) - (lo - 1) // 2)}):
            rec.put_seq(f'trunc{a}_{b}.{sk}', s.trunc((a, b)))


def reduce_case(rec: Rec, data: torch.Tensor, sizes: torch.Tensor, tag: str, with_head_last=True,
                store_inputs=True):
    if store_inputs:
        rec.put(f'{tag}.data', data, True)
        rec.put(f'{tag}.sizes', sizes, True)
    fns = ['sum', 'mean', 'prod', 'max', 'min', 'logsumexp'] + (['head', 'last'] if with_head_last else [])
    for fn in fns:
        rec.put(f'{tag}.{fn}', getattr(rua, 'segment_' + fn)(data, sizes), True)


def seg_case(rec: Rec, c: C, durations, fns=('sum', 'mean', 'max', 'min', 'logsumexp', 'last', 'prod')):
    d = C.new(durations)
    rec.put_seq('dur.C', d, True)
    for sk in 'CLPR':
        s = build(sk, c)
        for dk in 'CLPR':
            dd = build(dk, d)
            for fn in fns:
                out = s.seg(dd, getattr(rua, 'segment_' + fn))
                rec.put_seq(f'seg.{sk}.{dk}.{fn}', out, True)
    # segment_head works only on C / P sources in the reference
    for sk in 'CP':
        out = build(sk, c).seg(d, rua.segment_head)
        rec.put_seq(f'seg.{sk}.C.head', out, True)


def save(name, rec):
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **rec)
    print(f'{name}: {len(rec)} arrays, {os.path.getsize(path) / 1024:.1f} KiB')


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)

    # ---- small: ties in the lengths, fp32, H=2 ------------------------------------------------
    rec = Rec()
    lens = torch.tensor([3, 1, 4, 1, 2, 4, 3], dtype=torch.long)
    g = torch.Generator().manual_seed(1)
    c = C(data=torch.randn((int(lens.sum()), 2), generator=g), token_sizes=lens)
    layout_case(rec, c)
    save('small_f32', rec)

    # ---- featureless int64 payload (separate branch of cat_rev, select/rev.py:10-17) ------------
    rec = Rec()
    lens = torch.tensor([2, 5, 1, 3, 3], dtype=torch.long)
    c = C(data=torch.arange(100, 100 + int(lens.sum()), dtype=torch.long), token_sizes=lens)
    layout_case(rec, c)
    save('featureless_i64', rec)

    # ---- cfg1: BASELINE.json configs[0]: B=32, len~U[1,128], hidden 300 fp32 ------------------
    rec = Rec()
    g = torch.Generator().manual_seed(0)
    lens = torch.randint(1, 129, (32,), generator=g)
    data = torch.randn((int(lens.sum()), 300), generator=g)
    c = C(data=data, token_sizes=lens)
    layout_case(rec, c, fills=(0,))
    reduce_case(rec, data, lens, 'reduce', store_inputs=False)  # inputs = src.C.*
    save('cfg1_f32', rec)

    # ---- bf16 payload moves (2-byte rows, odd hidden) -----------------------------------------
    rec = Rec()
    g = torch.Generator().manual_seed(2)
    lens = torch.randint(1, 20, (9,), generator=g)
    data = torch.randn((int(lens.sum()), 5), generator=g).to(torch.bfloat16)
    layout_case(rec, C(data=data, token_sizes=lens), fills=(0, 2))
    save('small_bf16', rec)

    # ---- reductions: empty segments, NaN poisoning, trailing zeros, bf16 contract --------------
    rec = Rec()
    g = torch.Generator().manual_seed(3)
    sizes = torch.tensor([2, 0, 3, 0, 1, 6, 0], dtype=torch.long)
    data = torch.randn((int(sizes.sum()), 4), generator=g)
    reduce_case(rec, data, sizes, 'empty', with_head_last=False)
    rec.put('empty.last', rua.segment_last(data, sizes), True)  # empty segments wrap to the previous row
    nan = data.clone()
    nan[4, 1] = float('nan')
    reduce_case(rec, nan, sizes, 'nan', with_head_last=False)
    sizes1 = torch.tensor([5, 1, 9, 2], dtype=torch.long)
    data1 = torch.randn((int(sizes1.sum()),), generator=g)
    reduce_case(rec, data1, sizes1, 'flat')
    data64 = torch.randn((int(sizes1.sum()), 3), generator=g, dtype=torch.float64)
    reduce_case(rec, data64, sizes1, 'f64')
    # long segments (chunk-spanning in the CUDA kernel) incl. one empty and one of length 1
    sizes2 = torch.tensor([700, 1, 0, 33, 257, 128, 5], dtype=torch.long)
    data2 = torch.randn((int(sizes2.sum()), 7), generator=g)
    reduce_case(rec, data2, sizes2, 'long', with_head_last=False)
    # bf16 contract: the reference evaluated on the same bf16 values upcast to fp32
    bf = torch.randn((int(sizes2.sum()), 8), generator=g).to(torch.bfloat16)
    rec.put('bf16.data', bf, True)
    rec.put('bf16.sizes', sizes2, True)
    for fn in ['sum', 'mean', 'max', 'min', 'logsumexp', 'prod']:
        out32 = getattr(rua, 'segment_' + fn)(bf.float(), sizes2)
        rec.put(f'bf16.{fn}.f32', out32, True)
        rec.put(f'bf16.{fn}.rounded', out32.to(torch.bfloat16), True)
        rec.put(f'bf16.{fn}.native', getattr(rua, 'segment_' + fn)(bf, sizes2), True)
    # offsets clamp quirk (layout/cat.py:81): trailing zero lengths
    cz = C(data=torch.arange(3.), token_sizes=torch.tensor([2, 1, 0]))
    rec.put('clamp.offsets', cz.offsets(), True)
    save('reduce_edge', rec)

    # ---- .seg(duration, fn): 4 layouts x 4 duration layouts x reducers --------------------------
    rec = Rec()
    g = torch.Generator().manual_seed(4)
    lens = torch.tensor([5, 2, 7, 1, 4], dtype=torch.long)
    data = torch.randn((int(lens.sum()), 3), generator=g)
    durations = []
    for n in lens.tolist():
        cuts = torch.unique(torch.randint(n, (n,), generator=g), sorted=False, return_counts=True)[1]
        durations.append(cuts)
    c = C(data=data, token_sizes=lens)
    rec.put_seq('src.C', c, True)
    seg_case(rec, c, durations)
    save('seg_f32', rec)

    # ---- tie-order hazard: the reference's non-stable sort at n=2000 --------------------------
    rec = Rec()
    g = torch.Generator().manual_seed(5)
    lens = torch.randint(1, 65, (2000,), generator=g)
    c = C(data=torch.arange(int(lens.sum()), dtype=torch.int32), token_sizes=lens)
    p = c.pack()
    rec.put_seq('src.C', c, True)
    rec.put('pack.batch_sizes', p.batch_sizes, True)
    rec.put('pack.sorted_indices', p.sorted_indices, True)
    rec.put('pack.unsorted_indices', p.unsorted_indices, True)
    rec.put('pack.data', p.data)
    rec.put('pack.roll3.data', p.roll(3).data)
    rec.put('pack.rev.data', p.rev().data)
    rec.put('pack.last', p.last())
    rec.put('pack.left.data', p.left(-1).data)
    save('ties_2000', rec)


def compose_case():
    """compose(list of ragged batches in mixed layouts) -> P; stored in canonical (cat) form, plus split()"""
    rec = Rec()
    g = torch.Generator().manual_seed(6)
    batches = []
    for b, kind in enumerate('CLPR'):
        lens = torch.randint(1, 7, (3 + b,), generator=g)
        data = torch.randn((int(lens.sum()), 3), generator=g)
        rec.put(f'in{b}.data', data, True)
        rec.put(f'in{b}.token_sizes', lens, True)
        batches.append(build(kind, C(data=data, token_sizes=lens)))
    out = rua.compose(batches)
    rec.put_seq('out', out, True)
    cat = out.cat()
    rec.put_seq('out.cat', cat, True)
    for k, piece in enumerate(batches[2].split()):
        rec.put(f'split2.{k}', piece, True)
    for k, piece in enumerate(batches[1].split()):
        rec.put(f'split1.{k}', piece, True)
    rec.put('tolist3', np.asarray([len(x) for x in batches[3].tolist()], dtype=np.int64), True)
    save('compose_f32', rec)


def scatter_case():
    """scatter_{sum,mean,prod,max,min,logsumexp}: forward and gradients, include_self on/off, with rows of
    `tensor` that no index points at."""
    rec = Rec()
    g = torch.Generator().manual_seed(8)
    m, k, h = 9, 40, 3
    index = torch.randint(0, m - 2, (k,), generator=g)      # rows m-2, m-1 stay untouched
    index[index == 3] = 4                                   # and so does row 3
    tensor = torch.randn((m, h), generator=g)
    source = torch.randn((k, h), generator=g)
    weight = torch.randn((m, h), generator=g)
    rec.put('index', index, True)
    rec.put('tensor', tensor, True)
    rec.put('source', source, True)
    rec.put('weight', weight, True)
    for op in ('sum', 'mean', 'prod', 'max', 'min', 'logsumexp'):
        for inc in (False, True):
            t = tensor.clone().requires_grad_(True)
            s = (source * 0.3 + 1.0 if op == 'prod' else source).clone().requires_grad_(True)
            out = getattr(rua, 'scatter_' + op)(t, index, s, include_self=inc)
            gt, gs = torch.autograd.grad((out * weight)[torch.isfinite(out)].sum(), [t, s], allow_unused=True)
            tag = f'{op}.{int(inc)}'
            rec.put(tag + '.out', out, True)
            rec.put(tag + '.grad_tensor', gt if gt is not None else torch.zeros_like(t), True)
            rec.put(tag + '.grad_source', gs if gs is not None else torch.zeros_like(s), True)
    save('scatter_f32', rec)


def api_surface():
    """public names of the reference package and the methods bound on the four layouts"""
    skip = {'Any', 'List', 'Tuple', 'Union', 'Number', 'NamedTuple', 'Tensor', 'PackedSequence', 'torch', 'Key',
            'Value'}
    names = sorted(n for n in dir(rua) if not n.startswith('_') and n not in skip
                   and not isinstance(getattr(rua, n), type(os)))
    with open(os.path.join(OUT, 'api_names.txt'), 'w') as fp:
        fp.write('\n'.join(names) + '\n')
    lines = []
    for k, cls in KINDS.items():
        for m in sorted(dir(cls)):
            if m.startswith('_') or m in ('count', 'index'):
                continue
            if k == 'P' and m in dir(tuple):
                continue
            lines.append(f'{k} {m}')
    with open(os.path.join(OUT, 'api_methods.txt'), 'w') as fp:
        fp.write('\n'.join(lines) + '\n')
    print(f'api: {len(names)} names, {len(lines)} methods')


if __name__ == '__main__':
    main()
    compose_case()
    scatter_case()
    api_surface()
