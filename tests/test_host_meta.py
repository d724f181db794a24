"""CPU: the host side of the list constructors (torchrua_b200._native.host_metadata: offsets, N, T, stable descending
order, batch_sizes, their prefix sums -- computed with numpy from the tensors' shapes, no device sync) against the
oracle's closed forms and torch's own pack_sequence, plus the list planner's accept / reject rules."""
import numpy as np
import pytest
import torch
from torch.nn.utils.rnn import pack_sequence

from oracle import rua_oracle as ora
from torchrua_b200 import _native


def cases():
    g = np.random.default_rng(0)
    yield [3, 1, 5, 5, 2]
    yield [1]
    yield [0, 0, 4, 0]
    yield [7] * 9
    yield g.integers(0, 40, 300).tolist()
    yield np.minimum(g.zipf(1.5, 500), 300).tolist()


@pytest.mark.parametrize('lengths', list(cases()))
def test_host_metadata_matches_oracle(lengths):
    b, n, t, parts, bs = _native.host_metadata(lengths, True)
    lens = np.asarray(lengths, dtype=np.int64)
    assert (b, n, t) == (lens.size, int(lens.sum()), int(lens.max()))
    ln, off, srt, uns, bs2, poff = parts
    assert all(a.dtype == np.int64 for a in parts)
    assert np.array_equal(ln, lens) and np.array_equal(off, np.concatenate(([0], np.cumsum(lens))))
    o_bs, o_srt, o_uns = ora.pack_meta(lens)                         # stable descending order, like the device sort
    assert np.array_equal(srt, o_srt) and np.array_equal(uns, o_uns)
    assert np.array_equal(bs, o_bs) and np.array_equal(bs2, o_bs)
    assert np.array_equal(poff, np.concatenate(([0], np.cumsum(o_bs))))
    assert np.array_equal(ora.lengths_from_pack(bs, uns), lens)       # and back
    if lens.min() > 0:                                                # torch refuses empty sequences
        ref = pack_sequence([torch.zeros(k) for k in lengths], enforce_sorted=False)
        assert np.array_equal(bs, ref.batch_sizes.numpy())
        assert np.array_equal(lens[srt], lens[ref.sorted_indices.numpy()])


def test_host_metadata_without_pack_side():
    b, n, t, parts, bs = _native.host_metadata([2, 0, 3], False)
    assert (b, n, t) == (3, 5, 3) and len(parts) == 2 and bs is None
    assert parts[1].tolist() == [0, 2, 2, 5]


def test_list_plan_rejects_what_the_kernel_cannot_take():
    a, b = torch.zeros(2, 3), torch.zeros(1, 3)
    assert _native.list_plan([]) is None
    assert _native.list_plan([a, b]) is None                                   # CPU tensors: ATen path
    assert _native.list_plan([a, 'x']) is None
    assert _native.list_plan([torch.zeros(())]) is None
