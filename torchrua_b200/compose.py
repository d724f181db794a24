"""compose(list of ragged batches) -> one PackedSequence of sequences-of-sequences; mirror of
torchrua/compose.py:9-33.  Outside the (a)-(e) hot path: a composition of
idx / cat / pack / invert_permutation / gather, all of which are native here."""
from typing import List

import torch

from torchrua_b200.layout import C, P, Z
from torchrua_b200.utils import invert_permutation


def compose(sequences: List[Z]) -> P:
    offset, data, indices, token_sizes = 0, [], [], []
    for sequence in sequences:
        raw = sequence.raw()
        data.append(raw)
        idx, sizes = sequence.idx().cat()
        indices.append(idx + offset)
        token_sizes.append(sizes)
        offset += raw.size()[0]

    token_sizes = C.new(token_sizes)
    unsorted_indices = token_sizes.idx().pack().data

    indices = C(data=torch.cat(indices, dim=0), token_sizes=token_sizes.data).pack()
    unsorted_indices = indices.unsorted_indices[unsorted_indices]
    indices = indices._replace(
        sorted_indices=invert_permutation(unsorted_indices),
        unsorted_indices=unsorted_indices,
    )
    return torch.cat(data, dim=0)[indices]
