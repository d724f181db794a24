"""Stand-in for the reference's (unavailable) test helper package ``torchnyan``: the ten names its tests import
(SURVEY.md section 4), so that tests/test_gpu_reference_style.py can read like the reference's own tests."""
import torch
from hypothesis import strategies as st

BATCH_SIZE = 12
TOKEN_SIZE = 24
FEATURE_DIM = 8
TINY_BATCH_SIZE = 5
TINY_TOKEN_SIZE = 11

device = torch.device('cuda') if torch.cuda.is_available() else torch.device('cpu')


def sizes(*bounds):
    """sizes(A) -> one int in [1, A]; sizes(A, B) -> list (1..A long) of ints in [1, B]; sizes(A, B, C) -> nested."""
    *outer, last = bounds
    strategy = st.integers(min_value=1, max_value=last)
    for n in reversed(outer):
        strategy = st.lists(strategy, min_size=1, max_size=n)
    return strategy


def assert_close(actual, expected, rtol=1e-4, atol=1e-5, **kwargs):
    torch.testing.assert_close(actual, expected, rtol=rtol, atol=atol, check_stride=False, equal_nan=True)


def _is_pack(z):
    return hasattr(z, 'batch_sizes')


def assert_sequence_close(actual, expected, **kwargs):
    """field-by-field; PackedSequences are compared in canonical (cat) form plus batch_sizes, because the
    order of equal-length sequences inside a time step is arbitrary in the reference (non-stable CPU sort)."""
    assert type(actual) is type(expected), (type(actual), type(expected))
    if _is_pack(actual):
        assert torch.equal(actual.batch_sizes, expected.batch_sizes)
        actual, expected = actual.cat(), expected.cat()
    for a, e in zip(actual, expected):
        assert_close(a, e)


def assert_grad_close(actual, expected, inputs, **kwargs):
    if _is_pack(actual):
        actual, expected = actual.cat().data, expected.cat().data
    elif isinstance(actual, tuple):
        actual, expected = actual.data, expected.data
    cotangent = torch.randn_like(expected)
    ga = torch.autograd.grad(actual, inputs, cotangent, allow_unused=True, retain_graph=True)
    ge = torch.autograd.grad(expected, inputs, cotangent, allow_unused=True, retain_graph=True)
    for a, e in zip(ga, ge):
        if a is None or e is None:
            assert (a is None or not a.abs().any()) and (e is None or not e.abs().any())
        else:
            assert_close(a, e)
