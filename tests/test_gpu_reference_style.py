"""GPU parity, part 3: property tests written the way the reference tests its own API (hypothesis-drawn ragged
batches; expectations computed on the fly from torch's pack_sequence / pad_sequence and per-sequence torch
ops; forward AND gradients) -- see SURVEY.md section 4.  They run against ``import torchrua`` (the drop-in
alias of torchrua_b200), i.e. exactly the import line a user of the reference has."""
import pytest
import torch
from hypothesis import given, settings, strategies as st
from torch.nn import functional as F
from torch.nn.utils.rnn import pack_sequence, pad_sequence

from tests.nyan import (BATCH_SIZE, FEATURE_DIM, TINY_BATCH_SIZE, TINY_TOKEN_SIZE, TOKEN_SIZE, assert_close,
                        assert_grad_close, assert_sequence_close, device, sizes)

pytestmark = pytest.mark.gpu

import torchrua  # noqa: E402
from torchrua import C, L, P, R, Z  # noqa: E402

KINDS = Z.__args__
import os  # noqa: E402

# RUA_HYPOTHESIS_EXAMPLES=200 turns this file into a fuzzing session (the default keeps the suite under a minute)
EXAMPLES = int(os.environ.get('RUA_HYPOTHESIS_EXAMPLES', '12'))
COMMON = dict(deadline=None, max_examples=EXAMPLES, derandomize=EXAMPLES <= 12)


def expected_new(kind, tensors, padding_value=0):
    """the four layouts built WITHOUT this library (torch.cat / pack_sequence / pad_sequence / F.pad)."""
    lengths = torch.tensor([t.size()[0] for t in tensors], dtype=torch.long, device=tensors[0].device)
    if kind is C:
        return C(data=torch.cat(tensors, dim=0), token_sizes=lengths)
    if kind is P:
        return pack_sequence(tensors, enforce_sorted=False)
    if kind is L:
        return L(data=pad_sequence(tensors, batch_first=True, padding_value=padding_value), token_sizes=lengths)
    t = int(lengths.max())
    rows = [F.pad(x, pad=[0, 0] * (x.dim() - 1) + [t - x.size()[0], 0], value=padding_value) for x in tensors]
    return R(data=torch.stack(rows, dim=0), token_sizes=lengths)


def draw_inputs(token_sizes, dim):
    return [torch.randn((n, dim), device=device, requires_grad=True) for n in token_sizes]


@pytest.mark.parametrize('target', ['cat', 'left', 'pack', 'right'])
@settings(**COMMON)
@given(token_sizes=sizes(BATCH_SIZE, TOKEN_SIZE), dim=sizes(FEATURE_DIM),
       actual_kind=st.sampled_from(KINDS), expected_kind=st.sampled_from(KINDS))
def test_layout(target, token_sizes, dim, actual_kind, expected_kind):
    inputs = draw_inputs(token_sizes, dim)
    actual = getattr(actual_kind.new(inputs), target)()
    expected = getattr(expected_new(expected_kind, inputs), target)()
    assert_sequence_close(actual=actual, expected=expected)
    assert_grad_close(actual=actual, expected=expected, inputs=inputs)


@settings(**COMMON)
@given(token_sizes=sizes(BATCH_SIZE, TOKEN_SIZE), kind=st.sampled_from(KINDS),
       zero_one_dtype=st.sampled_from([
           (False, True, torch.bool), (-1, +2, torch.long),
           (torch.finfo(torch.float16).min, torch.finfo(torch.float16).max, torch.float16),
           (torch.finfo(torch.float32).min, torch.finfo(torch.float32).max, torch.float32),
           (torch.finfo(torch.float64).min, torch.finfo(torch.float64).max, torch.float64)]))
def test_mask(token_sizes, kind, zero_one_dtype):
    zero, one, dtype = zero_one_dtype
    inputs = [torch.randn((n,), device=device) for n in token_sizes]
    actual = kind.new(inputs).mask(zero=zero, one=one, dtype=dtype)
    expected = pad_sequence([torch.full((n,), fill_value=one, device=device, dtype=dtype) for n in token_sizes],
                            batch_first=True, padding_value=zero)
    assert_close(actual=actual, expected=expected)


@settings(**COMMON)
@given(data=st.data(), token_sizes=sizes(BATCH_SIZE, TOKEN_SIZE), dim=sizes(FEATURE_DIM), kind=st.sampled_from(KINDS))
def test_select(data, token_sizes, dim, kind):
    inputs = draw_inputs(token_sizes, dim)
    seq = kind.new(inputs)
    lo, hi = min(token_sizes), max(token_sizes)
    n = data.draw(st.integers(1, lo))
    shifts = data.draw(st.integers(-hi, hi))
    a = data.draw(st.integers(0, lo - 1))
    b = data.draw(st.integers(0, lo - 1 - a))
    cases = [
        (seq.head(n=n).cat(), C.new([x[:n] for x in inputs])),
        (seq.rev().cat(), C.new([x.flip(dims=[0]) for x in inputs])),
        (seq.roll(shifts=shifts).cat(), C.new([x.roll(shifts, dims=[0]) for x in inputs])),
        (seq.trunc((a, b)).cat(), C.new([x[a:x.size()[0] - b] for x in inputs])),
    ]
    for actual, expected in cases:
        assert_sequence_close(actual=actual, expected=expected)
        assert_grad_close(actual=actual.data, expected=expected.data, inputs=inputs)
    actual, expected = seq.last(), torch.stack([x[-1] for x in inputs], dim=0)
    assert_close(actual=actual, expected=expected)
    assert_grad_close(actual=actual, expected=expected, inputs=inputs)


DENSE = {
    'max': lambda x: x.max(dim=0).values, 'min': lambda x: x.min(dim=0).values, 'sum': lambda x: x.sum(dim=0),
    'mean': lambda x: x.mean(dim=0), 'prod': lambda x: x.prod(dim=0), 'logsumexp': lambda x: x.logsumexp(dim=0),
    'head': lambda x: x[0], 'last': lambda x: x[-1],
}


@pytest.mark.parametrize('name', sorted(DENSE))
@settings(**COMMON)
@given(token_sizes=sizes(BATCH_SIZE, TINY_TOKEN_SIZE), dim=sizes(FEATURE_DIM))
def test_segment_reduce(name, token_sizes, dim):
    inputs = draw_inputs(token_sizes, dim)
    expected = torch.stack([DENSE[name](x) for x in inputs], dim=0)
    tensor, segment_sizes = C.new(inputs)
    actual = getattr(torchrua, 'segment_' + name)(tensor, segment_sizes=segment_sizes)
    assert_close(actual=actual, expected=expected)
    assert_grad_close(actual=actual, expected=expected, inputs=inputs)


@pytest.mark.parametrize('name', ['max', 'min', 'sum', 'mean', 'prod', 'logsumexp'])
@settings(**COMMON)
@given(token_sizes=sizes(BATCH_SIZE, TINY_TOKEN_SIZE), dim=sizes(FEATURE_DIM), include_self=st.booleans())
def test_scatter_reduce(name, token_sizes, dim, include_self):
    inputs = draw_inputs(token_sizes, dim)
    index = torch.cat([torch.full((n,), fill_value=i, dtype=torch.long, device=device)
                       for i, n in enumerate(token_sizes)], dim=0)
    permutation = torch.randperm(sum(token_sizes), dtype=torch.long, device=device)
    tensor = torch.randn((len(token_sizes), dim), device=device)
    groups = [torch.cat([x, tensor[i:i + 1]], dim=0) if include_self else x for i, x in enumerate(inputs)]
    expected = torch.stack([DENSE[name](x) for x in groups], dim=0)
    source = torch.cat(inputs, dim=0)
    actual = getattr(torchrua, 'scatter_' + name)(tensor, index=index[permutation], source=source[permutation],
                                                  include_self=include_self)
    assert_close(actual=actual, expected=expected)
    assert_grad_close(actual=actual, expected=expected, inputs=inputs)


def raw_segment(rows, durations, fn):
    out = []
    for row, cuts in zip(rows, durations):
        start, pieces = 0, []
        for n in cuts.tolist():
            pieces.append(fn(row[start:start + n]))
            start += n
        out.append(torch.stack(pieces, dim=0))
    return out


@pytest.mark.parametrize('name', ['max', 'min', 'sum', 'mean', 'prod', 'logsumexp', 'last'])
@settings(**COMMON)
@given(token_sizes=sizes(BATCH_SIZE, TOKEN_SIZE), dim=sizes(FEATURE_DIM),
       seq_kind=st.sampled_from(KINDS), dur_kind=st.sampled_from(KINDS))
def test_seg(name, token_sizes, dim, seq_kind, dur_kind):
    inputs = draw_inputs(token_sizes, dim)
    durations = [torch.unique(torch.randint(n, (n,), device=device), sorted=False, return_counts=True)[1]
                 for n in token_sizes]
    actual = seq_kind.new(inputs).seg(dur_kind.new(durations), getattr(torchrua, 'segment_' + name)).cat()
    expected = C.new(raw_segment(inputs, durations, DENSE[name]))
    assert_sequence_close(actual=actual, expected=expected)
    assert_grad_close(actual=actual.data, expected=expected.data, inputs=inputs)


@settings(deadline=None, max_examples=max(6, EXAMPLES // 4), derandomize=EXAMPLES <= 12)
@given(token_sizes_batch=sizes(TINY_BATCH_SIZE, TINY_BATCH_SIZE, TINY_TOKEN_SIZE), dim=sizes(FEATURE_DIM),
       hidden=sizes(FEATURE_DIM), kind=st.sampled_from(KINDS))
def test_compose_feeds_an_lstm(token_sizes_batch, dim, hidden, kind):
    """compose(batches) -> one PackedSequence -> cuDNN LSTM == running the LSTM on every batch separately."""
    rnn = torch.nn.LSTM(input_size=dim, hidden_size=hidden, bidirectional=True, bias=True).to(device)
    batches = [[torch.randn((n, dim), device=device) for n in token_sizes] for token_sizes in token_sizes_batch]
    separate = []
    for tensors in batches:
        _, (h, _) = rnn(pack_sequence(tensors, enforce_sorted=False))
        separate.append(h.transpose(0, 1).flatten(start_dim=1))        # (B_k, 2*hidden)
    composed = torchrua.compose([kind.new(tensors) for tensors in batches])
    _, (h, _) = rnn(composed)
    h = h.transpose(0, 1).flatten(start_dim=1)                          # (sum B_k, 2*hidden), sequence order
    # the composed PackedSequence lists the sequences batch-interleaved: its unsorted order is the outer pack
    counts = torch.tensor([len(t) for t in batches], dtype=torch.long, device=device)
    order = C(data=torch.arange(int(counts.sum()), device=device), token_sizes=counts).pack().data
    flat = torch.cat(separate, dim=0)
    assert_close(actual=h, expected=flat[order], rtol=1e-3, atol=1e-4)


@settings(**COMMON)
@given(token_sizes=sizes(BATCH_SIZE, TOKEN_SIZE), dim=sizes(FEATURE_DIM), kind=st.sampled_from(KINDS))
def test_split_round_trip(token_sizes, dim, kind):
    inputs = [torch.randn((n, dim), device=device) for n in token_sizes]
    for actual, expected in zip(kind.new(inputs).split(), inputs):
        assert_close(actual=actual, expected=expected)
