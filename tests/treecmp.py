"""Flatten / freeze / compare the result trees of tests/scenarios.py (test infrastructure)."""
import numpy as np
import torch


def flatten(tree, prefix=''):
    """nested dict / list / tuple of tensors and plain values -> {path: leaf}"""
    out = {}
    if isinstance(tree, dict):
        for k, v in tree.items():
            out.update(flatten(v, f'{prefix}/{k}' if prefix else str(k)))
    elif isinstance(tree, (list, tuple)):
        for k, v in enumerate(tree):
            out.update(flatten(v, f'{prefix}[{k}]'))
        out[prefix + '#len'] = len(tree)
    else:
        out[prefix] = tree
    return out


def freeze(leaf):
    """leaf -> numpy array storable in an .npz (bf16 as uint16 bits; the path carries a '#bf16' tag)."""
    if isinstance(leaf, torch.Tensor):
        t = leaf.detach().cpu().contiguous()
        if t.dtype == torch.bfloat16:
            return t.view(torch.uint16).numpy(), '#bf16'
        return t.numpy(), ''
    if leaf is None:
        return np.asarray('__none__'), ''
    return np.asarray(leaf), ''


def thaw(arr: np.ndarray, tag: str):
    if tag == '#bf16':
        return torch.from_numpy(arr.copy()).view(torch.bfloat16)
    if arr.dtype.kind in 'US':
        s = str(arr)
        return None if s == '__none__' else s
    if arr.dtype == np.bool_ and arr.shape == ():
        return bool(arr)
    if arr.shape == () and arr.dtype.kind in 'iu':
        return int(arr)
    return torch.from_numpy(arr.copy())


def _same_bits(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    if a.is_floating_point():
        return (a == b) | (a.isnan() & b.isnan())
    return a == b


def assert_exact(path, got, exp):
    if isinstance(exp, torch.Tensor):
        assert isinstance(got, torch.Tensor), f'{path}: expected a tensor, got {type(got).__name__}'
        assert got.shape == exp.shape, f'{path}: shape {tuple(got.shape)} != {tuple(exp.shape)}'
        assert got.dtype == exp.dtype, f'{path}: dtype {got.dtype} != {exp.dtype}'
        same = _same_bits(got, exp)
        assert bool(same.all()), f'{path}: {int((~same).sum())} of {same.numel()} entries differ'
    else:
        assert got == exp, f'{path}: {got!r} != {exp!r}'


def assert_close(path, got, exp, rtol, atol, scale=None):
    """|got - exp| <= rtol*|exp| + atol (+ rtol*scale); NaNs and infinities must coincide."""
    if not isinstance(exp, torch.Tensor) or not exp.is_floating_point():
        return assert_exact(path, got, exp)
    assert got.shape == exp.shape, f'{path}: shape {tuple(got.shape)} != {tuple(exp.shape)}'
    g, e = got.double(), exp.double()
    special = ~torch.isfinite(e)
    assert bool(_same_bits(g[special], e[special]).all()), f'{path}: non-finite entries differ'
    bound = rtol * e.abs() + atol
    if scale is not None:
        bound = bound + rtol * scale.double()
    bad = ((g - e).abs() > bound) & ~special
    assert not bool(bad.any()), (f'{path}: {int(bad.sum())} of {bad.numel()} entries outside tolerance '
                                 f'(max abs err {float((g - e).abs()[~special].max()):.3e}, rtol {rtol}, atol {atol})')


EXACT_REDUCE_LEAVES = ('max', 'min', 'head', 'last', 'abs_sum')


def compare(case_mode, got_tree, exp_tree, label=''):
    got, exp = flatten(got_tree), flatten(exp_tree)
    assert set(got) == set(exp), f'{label}: result trees differ in structure: {sorted(set(got) ^ set(exp))[:6]}'
    scale = exp.get('abs_sum')
    for path in sorted(exp):
        g, e = got[path], exp[path]
        where = f'{label}:{path}'
        if case_mode == 'exact' or not isinstance(e, torch.Tensor) or not e.is_floating_point():
            assert_exact(where, g, e)
        elif case_mode == 'close':
            assert_close(where, g, e, 1e-5, 1e-5)
        elif case_mode == 'reduce':
            op = path.split('/')[0].split('.')[0]
            if op == 'abs_sum':
                assert_close(where, g, e, 1e-12, 0.0)
            elif path in EXACT_REDUCE_LEAVES or (op in ('max', 'min') and not path.endswith('.grad')):
                assert_exact(where, g, e)
            elif e.dtype == torch.float64:
                assert_close(where, g, e, 1e-12, 1e-12)
            elif path.endswith('.grad'):
                assert_close(where, g, e, 1e-5, 1e-6)
            elif op in ('sum', 'mean') and scale is not None and scale.shape == e.shape:
                assert_close(where, g, e, 1e-5, 1e-7, scale=scale)       # fp32: rtol 1e-5 (+ 1e-5 * sum |x|)
            else:
                assert_close(where, g, e, 1e-5, 2e-5)
        elif case_mode == 'bf16':
            # reference evaluated in fp32 on the same bf16 values, rounded ONCE to bf16; ours: native bf16 in / out
            if path == 'abs_sum':
                continue
            ref16 = e.to(torch.bfloat16) if e.dtype == torch.float32 else e
            if path in EXACT_REDUCE_LEAVES:
                assert_exact(where, g, ref16)
            else:
                assert g.dtype == torch.bfloat16, f'{where}: dtype {g.dtype}'
                sc = scale if (scale is not None and scale.shape == e.shape) else None
                assert_close(where, g.float(), ref16.float(), 1e-2, 1e-6, scale=None if sc is None else sc * 1e-3)
        else:
            raise ValueError(case_mode)
