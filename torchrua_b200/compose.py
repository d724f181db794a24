"""compose(batches) -> one PackedSequence holding every sequence of every batch (mirror of
torchrua/compose.py:9-33; outside the (a)-(e) hot path).

Two nested packings, both done by the native kernels:
  * INNER: all sequences of all batches, taken as one ragged batch of flat storage rows, are packed
    (stable descending device sort, one row-map launch over the int64 row indices);
  * OUTER: the per-batch sequence counts are packed too, which yields the order in which a consumer
    (e.g. an RNN over "sequences of sequences") wants the sequences back.
The result's unsorted_indices is the composition of the two permutations.  The payload moves ONCE: the reference
concatenates every source (one read + one write of all payload bytes) and then gathers (another read + write); here
one multi-source launch (rua_gather_rows_multi) reads each row from the tensor it lives in and writes it to its place
in the packed result.
"""
from typing import List

import torch

from torchrua_b200 import _native
from torchrua_b200.layout import C, P, Z


def compose(sequences: List[Z]) -> P:
    payloads, row_ids, lengths, base = [], [], [], 0
    for z in sequences:
        flat = z.raw()
        cat_rows = z.idx().cat()                      # storage row of every token, sequence-major
        payloads.append(flat)
        row_ids.append(cat_rows.data + base)
        lengths.append(cat_rows.token_sizes)
        base += flat.size()[0]

    all_lengths = torch.cat(lengths, dim=0)
    counts = torch.tensor([n.size()[0] for n in lengths], dtype=torch.long, device=all_lengths.device)

    outer = C(data=torch.arange(all_lengths.size()[0], dtype=torch.long, device=all_lengths.device),
              token_sizes=counts).pack()
    inner = C(data=torch.cat(row_ids, dim=0), token_sizes=all_lengths).pack()

    unsorted_indices = inner.unsorted_indices[outer.data]
    return P(
        data=_native.gather_rows_multi(inner.data, payloads),      # no torch.cat pass: rows come straight from each source
        batch_sizes=inner.batch_sizes,
        sorted_indices=_native.invert_permutation(unsorted_indices),
        unsorted_indices=unsorted_indices,
    )
