"""Multi-GPU: shard a ragged batch BY SEQUENCE.

No op on the hot path mixes tokens of different sequences, so every rank runs the single-GPU kernels
on its own sequences and no payload byte crosses NVLink during compute.  The only exchanges are
  (1) an all-gather of the per-rank lengths (8 bytes per sequence) so that every rank can derive any
      global metadata (offsets, T, sorted order) redundantly with the K0 kernels, and
  (2) optionally a variable-size all-gather of the outputs back into global sequence order.
The reference has no multi-GPU support at all; this module is host-side plumbing over
``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests).
"""
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor


def balanced_partition(lengths: Tensor, world_size: int) -> List[Tensor]:
    """Length-balanced, deterministic assignment of sequences to ranks.

    Sequences are sorted by length (descending, stable) and dealt out in boustrophedon ("snake")
    order, so every rank receives the same number of sequences (+-1) and, for any smooth length
    distribution, the same number of tokens to within a fraction of the longest sequence.
    Returns, per rank, the ascending list of global sequence ids it owns.
    """
    lengths = lengths.detach().cpu().long()
    order = torch.sort(lengths, descending=True, stable=True)[1]
    k = torch.arange(order.numel())
    phase = k % (2 * world_size)
    rank_of_sorted = torch.where(phase < world_size, phase, 2 * world_size - 1 - phase)
    owner = torch.empty_like(rank_of_sorted)
    owner[order] = rank_of_sorted
    return [torch.nonzero(owner == r).view(-1) for r in range(world_size)]


def partition_imbalance(lengths: Tensor, parts: List[Tensor]) -> float:
    """max over ranks of tokens / mean tokens (1.0 = perfect)."""
    tokens = torch.tensor([int(lengths[p].sum()) for p in parts], dtype=torch.double)
    return float(tokens.max() / tokens.mean()) if tokens.numel() and float(tokens.mean()) > 0 else 1.0


def take_sequences(data: Tensor, token_sizes: Tensor, ids: Tensor) -> Tuple[Tensor, Tensor]:
    """host-side helper: the (data, token_sizes) of a CattedSequence restricted to sequences ``ids``."""
    off = torch.cumsum(token_sizes, 0) - token_sizes
    lens = token_sizes[ids]
    rows = torch.repeat_interleave(off[ids], lens) + (torch.arange(int(lens.sum())) -
                                                      torch.repeat_interleave(torch.cumsum(lens, 0) - lens, lens))
    return data[rows], lens


def all_gather_lengths(local_lengths: Tensor, group: Optional[dist.ProcessGroup] = None) -> List[Tensor]:
    """exchange (1): every rank learns every rank's lengths.  One small collective: the per-rank counts
    ride along in slot 0 of a fixed-size buffer, so a single all_gather suffices when the caller knows
    an upper bound ``cap`` on the per-rank batch; otherwise counts are exchanged first."""
    world = dist.get_world_size(group)
    n = torch.tensor([local_lengths.numel()], dtype=torch.long, device=local_lengths.device)
    counts = [torch.empty_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c) for c in counts]
    cap = max(counts)
    buf = torch.zeros(cap, dtype=torch.long, device=local_lengths.device)
    buf[:local_lengths.numel()] = local_lengths
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return [o[:c] for o, c in zip(out, counts)]


def all_gather_lengths_fixed(local_lengths: Tensor, cap: int, out: Tensor,
                             group: Optional[dist.ProcessGroup] = None, async_op: bool = False):
    """single-collective variant for steady-state loops: ``out`` is (world, cap+1); row r = [count_r,
    lengths_r..., 0 padding].  With ``async_op`` the collective is returned as a work handle so that it
    overlaps the local kernels (nothing on a rank's own data path waits for the other ranks' lengths)."""
    buf = torch.zeros(cap + 1, dtype=torch.long, device=local_lengths.device)
    buf[0] = local_lengths.numel()
    buf[1:1 + local_lengths.numel()] = local_lengths
    work = dist.all_gather_into_tensor(out.view(-1), buf, group=group, async_op=async_op)
    return work if async_op else out


def gather_rows_by_sequence(local_rows: Tensor, parts: List[Tensor], group: Optional[dist.ProcessGroup] = None
                            ) -> Tensor:
    """exchange (2) for per-sequence outputs (segment reductions, last(), head(1)): local_rows is
    (B_r, *) in the order of ``parts[rank]``; returns (B, *) in global sequence order on every rank."""
    world = dist.get_world_size(group)
    cap = max(p.numel() for p in parts)
    feat = tuple(local_rows.shape[1:])
    buf = local_rows.new_zeros((cap,) + feat)
    buf[:local_rows.shape[0]] = local_rows
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    total = sum(p.numel() for p in parts)
    result = local_rows.new_empty((total,) + feat)
    for r, ids in enumerate(parts):
        result[ids.to(result.device)] = out[r][:ids.numel()]
    return result


def gather_catted(local_data: Tensor, local_lengths: Tensor, parts: List[Tensor], global_lengths: Tensor,
                  group: Optional[dist.ProcessGroup] = None) -> Tensor:
    """exchange (2) for per-token outputs: all ranks receive the global (N, *) C data in original
    sequence order.  Variable sizes are padded to the largest shard for the collective."""
    world = dist.get_world_size(group)
    dev = local_data.device
    tokens = [int(global_lengths[p].sum()) for p in parts]
    cap = max(tokens)
    feat = tuple(local_data.shape[1:])
    buf = local_data.new_zeros((cap,) + feat)
    buf[:local_data.shape[0]] = local_data
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    gl = global_lengths.to(dev)
    goff = torch.cumsum(gl, 0) - gl
    result = local_data.new_empty((int(gl.sum()),) + feat)
    for r, ids in enumerate(parts):
        ids = ids.to(dev)
        lens = gl[ids]
        n = int(lens.sum())
        loc_off = torch.cumsum(lens, 0) - lens
        dst = torch.repeat_interleave(goff[ids], lens) + (torch.arange(n, device=dev) -
                                                          torch.repeat_interleave(loc_off, lens))
        result[dst] = out[r][:n]
    return result
