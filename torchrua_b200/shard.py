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


def micro_batch_cuts(lengths: Tensor, k: int) -> List[Tuple[int, int, int, int]]:
    """Cut a rank's sequences (host lengths, in local order) into at most ``k`` consecutive runs holding about the same number
    of tokens each: ``(first sequence, one past the last, first token, one past the last token)`` per run.  Every sequence
    lands in exactly one run; runs are never empty.  The output gather walks these micro-batches: the peer stores of one
    overlap the conversions of the next."""
    lengths = lengths.detach().cpu().long()
    b = lengths.numel()
    if b == 0:
        return []
    k = max(1, min(int(k), b))
    csum = torch.cumsum(lengths, 0)
    total = int(csum[-1])
    cuts, start = [], 0
    for j in range(k):
        if start >= b:
            break
        end = b if j == k - 1 else int(torch.searchsorted(csum, total * (j + 1) // k, right=True))
        end = min(max(end, start + 1), b)
        cuts.append((start, end, int(csum[start - 1]) if start else 0, int(csum[end - 1])))
        start = end
    if cuts and cuts[-1][1] < b:      # rounding left a tail: it joins the last run
        s0, _, t0, _ = cuts[-1]
        cuts[-1] = (s0, b, t0, total)
    return cuts


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


# ------------------------------------------------------------------------------------------------
# exchange (2), B200-native: every rank's kernel stores its rows straight into the output buffers of all
# ranks through NVLink peer mappings (K5, include/rua_b200.h).  No staging copy, no padding to the largest
# shard, no permutation pass afterwards.  Host side = plumbing only: window set-up and a stream-ordered fence.
# ------------------------------------------------------------------------------------------------
class _DeviceMemory:
    """adapter that lets torch wrap a raw device pointer without copying (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {'shape': (nbytes,), 'typestr': '|u1', 'data': (ptr, False), 'version': 2}


class PeerWindows:
    """one ``nbytes`` window per rank of ``group`` (all ranks on ONE node), each mapped into every process.

    ``ptrs[r]`` is the address of rank r's window in THIS process; ``local`` is this rank's window as a
    uint8 tensor.  Set-up synchronises (cudaMalloc + a host-side handle exchange); use it once and keep it.
    """

    def __init__(self, nbytes: int, group: Optional[dist.ProcessGroup] = None, device: Optional[torch.device] = None):
        import ctypes

        from torchrua_b200 import _lib
        self._lib = _lib.load()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > _lib.MAX_DESTINATIONS - 1:
            raise RuntimeError(f'torchrua_b200: at most {_lib.MAX_DESTINATIONS - 1} ranks per window group')
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.nbytes = (int(nbytes) + 255) // 256 * 256
        with torch.cuda.device(self.device):
            ptr, handle = ctypes.c_void_p(), ctypes.create_string_buffer(_lib.PEER_HANDLE_BYTES)
            _lib.check(self._lib.rua_peer_window_alloc(self.nbytes, ctypes.byref(ptr), handle), 'rua_peer_window_alloc')
            self._own = ptr.value
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle.raw), group=group)
            self.ptrs, self._opened = [], []
            for r, h in enumerate(handles):
                if r == self.rank:
                    self.ptrs.append(self._own)
                    continue
                peer = ctypes.c_void_p()
                _lib.check(self._lib.rua_peer_window_open(h, ctypes.byref(peer)), 'rua_peer_window_open')
                self.ptrs.append(peer.value)
                self._opened.append(peer.value)
            self.local = torch.as_tensor(_DeviceMemory(self._own, self.nbytes), device=self.device)
        self._flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.fence()

    def view(self, shape, dtype: torch.dtype, offset_bytes: int = 0) -> Tensor:
        n = 1
        for s in shape:
            n *= int(s)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        if offset_bytes % 256 or offset_bytes + nbytes > self.nbytes:
            raise RuntimeError('torchrua_b200: view does not fit the peer window (offsets are multiples of 256 bytes)')
        return self.local[offset_bytes:offset_bytes + nbytes].view(dtype).view(tuple(shape))

    def fence(self) -> None:
        """all ranks' earlier kernels (on their current streams) have finished before anything enqueued
        after the fence runs.  NCCL: an 4-byte all-reduce, stream-ordered, no host sync.  Other backends
        (gloo, used by the single-GPU two-process test): device sync + host barrier."""
        if dist.get_backend(self.group) == 'nccl':
            dist.all_reduce(self._flag, group=self.group)
        else:
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)

    def close(self) -> None:
        if getattr(self, '_own', None) is None:
            return
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)      # nobody is still storing into a window that is about to go away
        self.local = None
        for p in self._opened:
            self._lib.rua_peer_window_close(p)
        self._lib.rua_peer_window_free(self._own)
        self._own, self._opened, self.ptrs = None, [], []


def push_lengths(local_lengths: Tensor, cap: int, windows: PeerWindows, offset_bytes: int = 0) -> Tensor:
    """exchange (1) without a collective: row ``rank`` of the (world, cap + 1) int64 table that lives at
    ``offset_bytes`` of EVERY rank's window becomes [count, lengths...] -- ONE tiny kernel of plain peer stores
    (rua_scatter_rows_multi over 8-byte rows; the count slot is rewritten only when the batch size changes), no
    staging buffer, no rendezvous.  Entries beyond ``count`` are whatever was there before.  Returns this rank's
    table (a view of its window); call ``windows.fence()`` before reading other ranks' rows."""
    from torchrua_b200 import _lib, _native
    lib = _lib.load()
    dev = windows.device
    lens = local_lengths if local_lengths.dtype == torch.long else local_lengths.long()
    lens = lens.contiguous()
    n = lens.numel()
    if n > cap:
        raise RuntimeError(f'torchrua_b200: {n} local sequences exceed the table capacity {cap}')
    dsts = _pointer_array([p + offset_bytes for p in windows.ptrs])
    cache = windows.__dict__.setdefault('_push_cache', {})
    ent = cache.get((cap, offset_bytes))
    with torch.cuda.device(dev):
        if ent is None or ent[0] != n:
            base = windows.rank * (cap + 1)
            slots = torch.arange(base + 1, base + 1 + n, dtype=torch.long, device=dev)
            count = torch.tensor([n], dtype=torch.long, device=dev)
            first = torch.tensor([base], dtype=torch.long, device=dev)
            _lib.check(lib.rua_scatter_rows_multi(count.data_ptr(), first.data_ptr(), 1, 8, dsts, windows.world,
                                                  _native._stream()), 'rua_scatter_rows_multi')
            ent = cache[(cap, offset_bytes)] = (n, slots, count, first)
        if n > 0:
            _lib.check(lib.rua_scatter_rows_multi(lens.data_ptr(), ent[1].data_ptr(), n, 8, dsts, windows.world,
                                                  _native._stream()), 'rua_scatter_rows_multi')
    return windows.view((windows.world, cap + 1), torch.long, offset_bytes)


def _pointer_array(values):
    import ctypes
    return (ctypes.c_void_p * len(values))(*[v if v else None for v in values])


def global_offsets(global_lengths: Tensor, device: torch.device) -> Tuple[Tensor, int]:
    """(exclusive prefix sum of the GLOBAL lengths on `device`, total token count): where every sequence of the batch
    starts in the global C data.  Compute once per batch and hand it to the gathers of its micro-batches."""
    from torchrua_b200 import _native
    gl = global_lengths.to(device, non_blocking=True)
    goff, gstats = _native.scan(gl)
    n_total = int(global_lengths.sum()) if not global_lengths.is_cuda else int(_native.fetch(gstats)[0])
    return goff, n_total


def gather_catted_fused(z, parts: List[Tensor], global_lengths: Tensor, windows: PeerWindows,
                        offset_bytes: int = 0, local_copy: bool = False, fence: bool = True,
                        seq_ids: Optional[Tensor] = None, goff: Optional[Tuple[Tensor, int]] = None,
                        local_out: Optional[Tensor] = None):
    """fused conversion + exchange (2): ``z`` is this rank's shard in ANY layout (C, L, P or R; sequences in
    the order of ``parts[rank]``).  ONE kernel reads every local token once and stores it at its place in
    the global C data (original sequence order) inside the window of every rank -- and, with
    ``local_copy``, also into a contiguous local C buffer (what ``z.cat().data`` would be).
    Returns the global (N, *) tensor (a view of this rank's window; valid after the trailing fence),
    or (global, local) with ``local_copy``.  ``fence=False`` leaves both fences to the caller (several
    gathers into disjoint parts of the window can share one pair: ``windows.fence()`` before the first store
    and after the last).

    Micro-batches: ``z`` may hold only SOME of this rank's sequences -- ``seq_ids`` (device int64) are their GLOBAL
    sequence ids, ``goff`` = ``global_offsets(...)`` of the whole batch (computed once), ``local_out`` the slice of the
    local C buffer that receives the contiguous copy.  Launched on a side stream, the peer stores of micro-batch k
    then overlap the conversions of micro-batch k+1 (NVLink-bound and HBM-bound work share the GPU)."""
    import ctypes

    from torchrua_b200 import _lib, _native
    from torchrua_b200.core.cast import side_of
    lib = _lib.load()
    dev = windows.device
    rg = z._ragged()
    src = z.raw()
    if not src.is_contiguous():
        src = src.contiguous()
    feat = tuple(src.shape[1:])
    row_bytes = src.element_size()
    for f in feat:
        row_bytes *= f
    if goff is None:
        goff = global_offsets(global_lengths, dev)
    goff, n_total = goff
    ids = seq_ids if seq_ids is not None else parts[windows.rank].to(dev, non_blocking=True)
    base = goff[ids]                                              # first global row of each local sequence
    out = windows.view((n_total,) + feat, src.dtype, offset_bytes)
    dsts = [p + offset_bytes for p in windows.ptrs]
    bases = [base.data_ptr()] * windows.world
    local = None
    if local_out is not None:
        if not local_out.is_contiguous() or local_out.shape[0] != rg.N:
            raise RuntimeError('torchrua_b200: local_out must be a contiguous (tokens of z, *) buffer')
        local = local_out
        dsts.append(local.data_ptr())
        bases.append(None)
    elif local_copy:
        local = torch.empty((rg.N,) + feat, dtype=src.dtype, device=dev)
        dsts.append(local.data_ptr())
        bases.append(None)
    side = side_of(z, rg).c_struct()
    rgc = rg.c_struct()
    if fence:
        windows.fence()      # peers have finished reading what the previous gather left in their windows
    with torch.cuda.device(dev):
        _lib.check(lib.rua_row_map_multi(src.data_ptr(), row_bytes, ctypes.byref(rgc), ctypes.byref(side), rg.N,
                                         _pointer_array(dsts), _pointer_array(bases), len(dsts), _native._stream()),
                   'rua_row_map_multi')
    if fence:
        windows.fence()      # every rank's rows have landed everywhere
    base.record_stream(torch.cuda.current_stream(dev))
    return (out, local) if (local_copy or local_out is not None) else out


def gather_rows_fused(local_rows: Tensor, parts: List[Tensor], windows: PeerWindows, offset_bytes: int = 0,
                      fence: bool = True) -> Tensor:
    """exchange (2) for per-sequence outputs (segment reductions, last(), head(1)): row j of ``local_rows``
    belongs to global sequence ``parts[rank][j]``; one kernel stores it there in every rank's window."""
    from torchrua_b200 import _lib, _native
    lib = _lib.load()
    dev = windows.device
    rows = local_rows.detach()
    if not rows.is_contiguous():
        rows = rows.contiguous()
    feat = tuple(rows.shape[1:])
    row_bytes = rows.element_size()
    for f in feat:
        row_bytes *= f
    total = sum(p.numel() for p in parts)
    out = windows.view((total,) + feat, rows.dtype, offset_bytes)
    ids = parts[windows.rank].to(dev, non_blocking=True).contiguous()
    dsts = [p + offset_bytes for p in windows.ptrs]
    if fence:
        windows.fence()
    with torch.cuda.device(dev):
        _lib.check(lib.rua_scatter_rows_multi(rows.data_ptr(), ids.data_ptr(), rows.shape[0], row_bytes,
                                              _pointer_array(dsts), len(dsts), _native._stream()),
                   'rua_scatter_rows_multi')
    if fence:
        windows.fence()
    return out
