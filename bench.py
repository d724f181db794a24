#!/usr/bin/env python
"""bench.py -- tokens/s and achieved HBM GB/s of the ragged-sequence hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (per GPU): BASELINE.json configs[1] -- batch 4096, lengths U[1,512], hidden 1024 bf16
(N ~ 1.06 M tokens, 2.17 GB per ragged layout, 4.29 GB per padded layout: far larger than the 126 MB L2).
One STEP is one pass of the hot path over the batch:

    C -> P (pack)  ->  L (left pad)  ->  R (right pad)  ->  C (cat)  +  segment_sum  +  segment_max

i.e. 4 layout conversions (one rua_row_map launch each, plus their metadata kernels) and 2 segment
reductions, every step recomputing all metadata (the per-tensor metadata cache is cleared each step).
`value` = tokens through the whole step per second, summed over GPUs (weak scaling: per-GPU work fixed).

Prints ONE JSON line (see the keys at the bottom).

`--impl reference` times the UNMODIFIED reference (speedcell4/torchrua 0.5.1, staged under oracle/_ref by
oracle/make_ref.py) through its own public API on the box's host cores -- the same step, the same seeds -- at the
full configs[1] batch when that fits the time budget, else on the largest prefix of the batch that does (stated in
`cpu_baseline.sample`).  The C + OpenMP port of the path (oracle/rua_oracle.c) is reported beside it as `cpu_port`.
The repo arm additionally reports `reference_cuda`: the same reference driven on the SAME GPU (its stock ATen CUDA
path, CUDA events) -- the kernel-to-beat of SURVEY.md 8d -- and per-op records for BASELINE configs 3, 4 and 5.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 4096
MAX_LEN = 512
HIDDEN = 1024
METRIC = 'tokens/sec through pack/pad/cat + segment reduce (C->P->L->R->C + segment_sum + segment_max)'
WORKLOAD = 'configs[1]: pack/pad/cat conversions, batch 4096, lengths U[1,512], hidden 1024 bf16'


def workload_lengths(world: int):
    import torch
    g = torch.Generator().manual_seed(0)
    return torch.randint(1, MAX_LEN + 1, (BATCH * world,), generator=g)


STEP = 'C->P->L->R->C + segment_sum + segment_max'


def make_config(world: int, exchange: str = 'peer'):
    """the `config` object of BOTH arms (the driver compares them): what one step processes."""
    glens = workload_lengths(world)
    total = int(glens.sum())
    return {'workload': WORKLOAD, 'step': STEP, 'batch_per_gpu': BATCH, 'tokens_total': total, 'hidden': HIDDEN,
            'l2': 'inputs larger than L2: 2.17 GB per ragged layout, 4.29 GB per padded layout vs 126 MB',
            'metadata': 'recomputed every step (cache cleared)',
            'parallelism': f'sequence-sharded x{world}, length-balanced (snake), lengths exchange only'}


# --------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port, all host threads, on a bounded sample of the same workload
# --------------------------------------------------------------------------------------------------
def cpu_pipeline_factory(n_seq: int):
    import numpy as np
    import torch
    from oracle import c_oracle as co
    lens = workload_lengths(1)[:n_seq].numpy()
    n = int(lens.sum())
    g = torch.Generator().manual_seed(1)
    data = torch.randn((n, HIDDEN), generator=g).to(torch.bfloat16).view(torch.uint16).numpy()

    def step():
        bs, srt, uns = co.pack_meta(lens)
        p = co.move(data, 'C', 'P', lens, bs, uns)
        left = co.move(p, 'P', 'L', lens, bs, uns, fill=0)
        right = co.move(left, 'L', 'R', lens, fill=0)
        back = co.move(right, 'R', 'C', lens)
        s = co.segment_reduce(back, lens, 'sum', bf16=True)
        m = co.segment_reduce(back, lens, 'max', bf16=True)
        return back, s, m

    return step, n, data


def time_cpu(steps: int, warmup: int, budget_s: float):
    """returns (tokens_per_s, ms_per_step, n_seq, n_tokens, threads)"""
    from oracle import c_oracle as co
    co.build()
    co.use_all_cores()
    n_seq = BATCH
    while True:
        step, n_tok, _ = cpu_pipeline_factory(n_seq)
        t0 = time.perf_counter()
        step()
        one = time.perf_counter() - t0
        if one * (steps + warmup) <= budget_s or n_seq <= 32:
            break
        n_seq //= 2
    for _ in range(max(warmup - 1, 0)):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return n_tok * steps / dt, dt / steps * 1e3, n_seq, n_tok, co.num_threads()


# --------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# --------------------------------------------------------------------------------------------------
class Clocks:
    """nvidia-smi sampler.  Started before the warm-up (its start-up latency is longer than a short timed
    region); only samples whose timestamps fall inside the marked timed region are reported."""
    QUERY = ('timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index: int):
        self.index = index
        self.path = tempfile.mktemp(prefix='rua_clocks_', suffix='.csv')
        self.proc = None
        self.t_begin = self.t_end = None

    def start(self):
        try:
            self.fp = open(self.path, 'w')
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits',
                                          '-i', str(self.index), '-lms', '20'], stdout=self.fp,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def begin(self):
        self.t_begin = time.time()

    def end(self):
        self.t_end = time.time()

    def stop(self):
        import datetime
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.proc is None:
            return out
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fp.close()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        rows = []
        for line in open(self.path):
            f = [x.strip() for x in line.split(',')]
            if len(f) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], '%Y/%m/%d %H:%M:%S.%f').timestamp()
                rows.append((ts, float(f[2]), float(f[3]), [n for n, v in zip(names, f[6:10]) if v.lower().startswith('active')]))
            except ValueError:
                continue
        try:
            os.unlink(self.path)
        except OSError:
            pass
        if not rows:
            return out
        lo, hi = (self.t_begin or 0) - 0.005, (self.t_end or 1e18) + 0.005
        inside = [r for r in rows if lo <= r[0] <= hi]
        where = 'inside the timed region'
        if not inside:   # region shorter than the sampling period: take the sample closest to it
            mid = 0.5 * (lo + hi)
            inside = [min(rows, key=lambda r: abs(r[0] - mid))]
            where = 'nearest to the timed region (region shorter than the 20 ms sampling period)'
        reasons = sorted({n for r in inside for n in r[3]})
        out.update(sm_mhz=statistics.median(r[1] for r in inside), sm_max_mhz=max(r[2] for r in inside),
                   reasons=reasons, samples=len(inside), where=where)
        return out


# --------------------------------------------------------------------------------------------------
# reference arm: the UNMODIFIED reference (oracle/_ref) through its own public API
# --------------------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, 'oracle', '_ref')


def import_reference():
    """speedcell4/torchrua as staged by oracle/make_ref.py -- never the drop-in alias package at the repo root."""
    if not os.path.exists(os.path.join(REF_DIR, 'torchrua', '__init__.py')):
        return None
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or os.getcwd()) != ROOT]
    sys.path.insert(0, REF_DIR)
    import torchrua
    assert os.path.dirname(os.path.dirname(os.path.abspath(torchrua.__file__))) == REF_DIR, torchrua.__file__
    sys.path.insert(1, ROOT)
    return torchrua


def reference_pipeline_factory(ref, n_seq: int, device: str):
    import torch
    lens = workload_lengths(1)[:n_seq]
    n = int(lens.sum())
    g = torch.Generator().manual_seed(1)
    data = torch.randn((n, HIDDEN), generator=g).to(torch.bfloat16).to(device)
    lens = lens.to(device)

    def step():
        c = ref.C(data=data, token_sizes=lens)
        back = c.pack().left(0).right(0).cat()
        s = ref.segment_sum(back.data, back.token_sizes)
        m = ref.segment_max(back.data, back.token_sizes)
        return back, s, m

    return step, n, data


def time_reference_cpu(ref, steps: int, warmup: int, budget_s: float, n_seq: int = 0):
    """-> (tokens_per_s, ms_per_step, n_seq, n_tokens, threads).  n_seq = 0: the full batch if the whole run fits the
    budget (extrapolated from a 256-sequence probe: the reference is linear in tokens), else the largest power-of-two
    prefix that does."""
    import torch
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    if n_seq <= 0:
        probe, n_probe, _ = reference_pipeline_factory(ref, 256, 'cpu')
        probe()
        t0 = time.perf_counter()
        probe()
        per_token = (time.perf_counter() - t0) / n_probe
        n_seq = BATCH
        total_tokens = int(workload_lengths(1).sum())
        while n_seq > 64 and per_token * total_tokens * (n_seq / BATCH) * (steps + warmup) > budget_s:
            n_seq //= 2
    step, n_tok, data = reference_pipeline_factory(ref, n_seq, 'cpu')
    for _ in range(warmup):
        out = step()
    if warmup:
        assert torch.equal(out[0].data, data), 'reference round trip is not the identity'
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return n_tok * steps / dt, dt / steps * 1e3, n_seq, n_tok, threads


def time_reference_cuda(ref, steps: int, warmup: int, n_seq: int):
    """the reference's stock ATen CUDA path on the same GPU, CUDA events around `steps` steps."""
    import torch
    torch.cuda.set_device(0)
    step, n_tok, data = reference_pipeline_factory(ref, n_seq or BATCH, 'cuda')
    for _ in range(max(warmup, 1)):
        out = step()
    torch.cuda.synchronize()
    assert torch.equal(out[0].data, data), 'reference round trip is not the identity'
    del out
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    # the same ops one by one (CUDA events around each public-API call of the reference, median of 5 after 2 warm-ups):
    # the per-op kernel-to-beat of SURVEY.md 8d
    lens = workload_lengths(1)[:n_seq or BATCH].to('cuda')
    c = ref.C(data=data, token_sizes=lens)
    p = c.pack()
    left = p.left(0)
    right = left.right(0)
    ops = {'C->P': lambda: c.pack(), 'P->L': lambda: p.left(0), 'L->R': lambda: left.right(0), 'R->C': lambda: right.cat(),
           'segment_sum': lambda: ref.segment_sum(data, lens), 'segment_max': lambda: ref.segment_max(data, lens)}
    per_op = {}
    for name, fn in ops.items():
        times = []
        for it in range(7):
            torch.cuda.synchronize()
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            if it >= 2:
                times.append(a.elapsed_time(b))
        per_op[name] = round(statistics.median(times), 4)
    return n_tok / (ms * 1e-3), ms, n_seq or BATCH, n_tok, per_op


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    ref = import_reference()
    port = ref_ops = None
    if args.ref_device == 'cuda':
        assert ref is not None, 'oracle/_ref is not staged'
        tps, ms, n_seq, n_tok, ref_ops = time_reference_cuda(ref, args.steps, args.warmup, args.ref_seqs)
        kind, threads, device = 'reference', 0, 'the same GPU (stock ATen CUDA kernels), CUDA events'
    elif ref is not None:
        tps, ms, n_seq, n_tok, threads = time_reference_cpu(ref, args.steps, max(args.warmup, 1), args.ref_budget,
                                                            args.ref_seqs)
        kind, device = 'reference', 'host CPU (no GPU work)'
    else:   # the staged reference did not travel: fall back to the C port of the same path, and say so
        tps, ms, n_seq, n_tok, threads = time_cpu(args.steps, max(args.warmup, 1), budget_s=150.0)
        kind, device = 'port', 'host CPU (no GPU work)'
    if ref is not None and args.ref_device == 'cpu' and not args.no_port:
        try:
            ptps, pms, pseq, ptok, pthreads = time_cpu(steps=3, warmup=1, budget_s=30.0)
            port = {'value': ptps, 'unit': 'tokens/s', 'cores': pthreads, 'kind': 'port', 'ms_per_step': pms,
                    'sample': sample_text(pseq, ptok),
                    'what': 'oracle/rua_oracle.c: C + OpenMP restatement of the same path, all host threads'}
        except Exception as e:   # the port is a second opinion, never the headline
            port = {'unavailable': f'{type(e).__name__}: {e}'}
    what = ('UNMODIFIED speedcell4/torchrua 0.5.1 (oracle/_ref) through its public API: C(...).pack().left(0).right(0).cat(), '
            'segment_sum, segment_max' if kind == 'reference' else
            'oracle/rua_oracle.c: C restatement of the reference path, OpenMP over all host threads (oracle/_ref missing)')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': tps, 'unit': 'tokens/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
        'config': make_config(args.gpus),
        'device': device,
        'cpu_baseline': {'value': tps, 'unit': 'tokens/s', 'cores': threads, 'kind': kind, 'ms_per_step': ms,
                         'sample': sample_text(n_seq, n_tok), 'full_batch': n_seq == BATCH, 'what': what},
        'cpu_port': port,
        'e2e': {'value': tps, 'unit': 'tokens/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    if ref_ops is not None:
        line['per_op_ms'] = ref_ops
    emit(line)


def sample_text(n_seq: int, n_tok: int) -> str:
    if n_seq == BATCH:
        return f'the full batch: all {BATCH} sequences ({n_tok} tokens, hidden {HIDDEN} bf16) per step'
    return f'first {n_seq} of the {BATCH} sequences ({n_tok} tokens, hidden {HIDDEN} bf16) per step'


def reference_subprocess(extra, timeout_s: float):
    """run `bench.py --impl reference ...` in a fresh interpreter (the reference and the package under test both
    patch torch.Tensor / PackedSequence process-wide and cannot share one) and parse its JSON line."""
    env = dict(os.environ)
    for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK', 'MASTER_ADDR', 'MASTER_PORT'):
        env.pop(k, None)
    try:
        out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--no-port'] + extra,
                             stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, timeout=timeout_s, cwd=ROOT)
        lines = [ln for ln in out.stdout.decode().splitlines() if ln.startswith('{')]
        if out.returncode != 0 or not lines:
            return {'unavailable': f'rc {out.returncode}: {out.stderr.decode()[-300:]}'}
        return json.loads(lines[-1])
    except Exception as e:
        return {'unavailable': f'{type(e).__name__}: {e}'}


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local: int):
    """Pin this rank's host threads to the CPUs NVML reports as local to its GPU BEFORE any pinned buffer is
    allocated (first touch then places the staging memory on the GPU's NUMA node).  Only matters for the e2e leg of
    multi-GPU runs, where all ranks stream through host memory at once.  Best effort: returns what was done."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:   # NVML does not honour CUDA_VISIBLE_DEVICES: address the device by its PCI id
            import torch
            pr = torch.cuda.get_device_properties(local)
            handle = pynvml.nvmlDeviceGetHandleByPciBusId(f'{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0')
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(local)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return f'bound to {len(os.sched_getaffinity(0))} CPUs local to GPU {local}'
    except Exception as e:   # no NVML / no permission: keep whatever the launcher set
        return f'unbound ({type(e).__name__})'


def run_ours(args):
    import torch
    import torch.distributed as dist

    import torchrua_b200 as rua
    from torchrua_b200 import _lib, _native, shard

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus and world > 1:
        raise SystemExit(f'--gpus {args.gpus} but WORLD_SIZE={world}')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else 'not bound (single GPU: the cpu_baseline leg uses every host core)'
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.load()

    # ---- synthetic batch, sharded by sequence with length-balanced partitioning --------------------
    glens = workload_lengths(world)
    parts = shard.balanced_partition(glens, world)
    lens_host = glens[parts[rank]].contiguous()
    n_tok = int(lens_host.sum())
    total_tokens = int(glens.sum())
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    data = torch.randn((n_tok, HIDDEN), generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    lens = lens_host.to(dev)
    cap = max(p.numel() for p in parts)
    gathered = torch.empty((world, cap + 1), dtype=torch.long, device=dev) if world > 1 else None
    meta_windows = shard.PeerWindows(world * (cap + 1) * 8) if world > 1 and args.exchange == 'peer' else None

    def step(d, ln):
        _native._CACHE.clear()                      # no metadata survives from one step to the next
        work = None
        c = rua.C(data=d, token_sizes=ln)
        p = c.pack()
        if meta_windows is not None:
            # the only exchange on the path: 8 bytes per sequence of metadata.  One tiny kernel stores this rank's
            # [count, lengths] row into the (world, cap+1) table of EVERY rank through the NVLink peer windows:
            # no collective call (an NCCL all-gather costs ~0.15 ms of host time per step here, measured), no
            # rendezvous between ranks.  A consumer of the table fences first (drain() below, once per timed region).
            shard.push_lengths(ln, cap, meta_windows)
        elif world > 1 and args.exchange == 'nccl':  # portable path: the same exchange as an async NCCL all-gather
            work = shard.all_gather_lengths_fixed(ln, cap, gathered, async_op=True)
        left = p.left(0)
        right = left.right(0)
        back = right.cat()
        s = rua.segment_sum(back.data, back.token_sizes)
        m = rua.segment_max(back.data, back.token_sizes)
        if work is not None:
            work.wait()                             # stream-ordered: the step is not done before the exchange is
        return back, s, m

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up, then the device-timed region -----------------------------------------------------
    clocks = Clocks(local)
    clocks.start()
    for _ in range(max(args.warmup, 3)):
        out = step(data, lens)
    assert torch.equal(out[0].data, data), 'round trip C->P->L->R->C is not the identity'
    if meta_windows is not None:                    # the exchanged table holds every rank's lengths
        meta_windows.fence()
        table = meta_windows.view((world, cap + 1), torch.long).cpu()
        for r in range(world):
            k = parts[r].numel()
            assert int(table[r, 0]) == k and torch.equal(table[r, 1:1 + k], glens[parts[r]]), 'lengths exchange'
    del out
    barrier()
    clocks.begin()
    _native.PROFILE = []
    launches0 = lib.rua_launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        step(data, lens)
    if meta_windows is not None:
        meta_windows.fence()                        # every rank's lengths have landed everywhere: part of the timed region
    t1.record()
    barrier()
    clocks.end()
    launches = lib.rua_launch_count() - launches0
    prof, _native.PROFILE = _native.PROFILE, None
    clk = clocks.stop()
    ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.double, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms)
    value = total_tokens * args.steps / (ms_total * 1e-3)

    # ---- per-kernel roofline from the events recorded inside the timed region ---------------------
    per = {}
    for name, a, b, nbytes in prof:
        e = per.setdefault(name, [0.0, 0, 0])
        e[0] += a.elapsed_time(b)
        e[1] += nbytes
        e[2] += 1
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'measured (MEASURED_PEAKS.json hbm_gbs)' if 'hbm_gbs' in peaks else 'fallback 6.65 TB/s (B200_PROFILING.md)'
    # dram__bytes_read.sum + dram__bytes_write.sum per launch of the same kernel from the committed `ncu --set full`
    # capture of this command (profiles/roofline_traffic.json names the capture it came from); never measured live
    traffic = traffic_src = None
    try:
        tr = json.load(open(os.path.join(ROOT, 'profiles', 'roofline_traffic.json')))
        traffic, traffic_src = tr.get('row_map_bytes_per_launch'), tr.get('source')
    except Exception:
        pass
    rm = per.get('row_map', [1.0, 0, 1])
    achieved = rm[1] / (rm[0] * 1e-3) / 1e9
    roofline = {'bound': 'hbm', 'kernel': 'row_map_kernel<V256> (256-bit vectors; 4 launches per step: C->P, P->L, L->R, R->C)',
                'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
                'traffic_source': traffic_src, 'peak_source': peak_src, 'frac_of_nameplate_8000': achieved / 8000.0,
                'algorithmic_bytes_per_launch': rm[1] / max(rm[2], 1), 'avg_launch_ms': rm[0] / max(rm[2], 1),
                'share_of_step': rm[0] / ms_total}
    sr = per.get('segment_reduce')
    kernels = {'row_map': {'gbs': achieved, 'ms_per_launch': rm[0] / max(rm[2], 1), 'launches': rm[2]}}
    if sr:
        kernels['segment_reduce'] = {'gbs': sr[1] / (sr[0] * 1e-3) / 1e9, 'ms_per_launch': sr[0] / sr[2],
                                     'launches': sr[2], 'share_of_step': sr[0] / ms_total}

    if args.gather_only:     # diagnostic: only the output-gather leg (fused path), one short JSON line
        g = time_output_gather(args, step, data, lens, lens_host, glens, parts, rank, world, dev, total_tokens, barrier)
        if rank == 0:
            emit({'gather_only': g, 'n_gpus': world, 'ms_per_step_no_gather': ms_total / args.steps,
                  'env': {k: v for k, v in os.environ.items() if k.startswith('RUA_')}})
        dist.destroy_process_group()
        return

    # ---- end to end through the public API with HOST buffers (H2D + D2H inside the timed region) --
    e2e_steps = max(4, min(args.steps, 20))     # pipeline fill and drain (one un-overlapped copy each way) are inside the timed region
    h_data = torch.empty((n_tok, HIDDEN), dtype=torch.bfloat16, pin_memory=True)
    h_data.copy_(data)
    h_lens = lens_host.pin_memory()
    h_back = torch.empty((n_tok, HIDDEN), dtype=torch.bfloat16, pin_memory=True)
    h_sum = torch.empty((lens_host.numel(), HIDDEN), dtype=torch.bfloat16, pin_memory=True)
    h_max = torch.empty((lens_host.numel(), HIDDEN), dtype=torch.bfloat16, pin_memory=True)  # (empty_like drops pinning)

    # Three streams: copy-in, compute, copy-out.  Every step still moves its own inputs host->device and its
    # own results device->host inside the timed region; the copies of neighbouring steps overlap with
    # compute and with each other (PCIe is full duplex), which is how a serving loop would drive this.
    s_in, s_comp, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    d_in = [torch.empty_like(data) for _ in range(2)]
    d_len = [torch.empty_like(lens) for _ in range(2)]
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    for ev in ev_free:
        ev.record(torch.cuda.current_stream())

    def e2e_step(k):
        b = k & 1
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_free[b])                 # buffer b is no longer read by step k-2
            d_in[b].copy_(h_data, non_blocking=True)
            d_len[b].copy_(h_lens, non_blocking=True)
            ev_in[b].record(s_in)
        with torch.cuda.stream(s_comp):
            s_comp.wait_event(ev_in[b])
            back, s, m = step(d_in[b], d_len[b])
            ev_free[b].record(s_comp)
            done = torch.cuda.Event()
            done.record(s_comp)
        with torch.cuda.stream(s_out):
            s_out.wait_event(done)
            for host_buf, t in ((h_back, back.data), (h_sum, s), (h_max, m)):
                host_buf.copy_(t, non_blocking=True)
                t.record_stream(s_out)

    def e2e_sync():
        for st_ in (s_in, s_comp, s_out):
            st_.synchronize()
        barrier()

    barrier()
    e2e_step(0)
    e2e_sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s_in)                                      # the first timed copy-in starts here
    for k in range(e2e_steps):
        e2e_step(k)
    s_out.wait_stream(s_comp)
    s_out.wait_stream(s_in)
    e1.record(s_out)                                     # the last result has landed on the host
    e2e_sync()
    assert torch.equal(h_back, h_data)
    ems = torch.tensor([e0.elapsed_time(e1)], dtype=torch.double, device=dev)
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    e2e_value = total_tokens * e2e_steps / (float(ems) * 1e-3)
    h2d = h_data.numel() * 2 + h_lens.numel() * 8
    d2h = h_back.numel() * 2 + h_sum.numel() * 2 * 2

    # ---- output gather (N > 1 only): every rank ends up with the GLOBAL results ----------------------
    gather = None
    if world > 1 and not args.no_gather:
        gather = time_output_gather(args, step, data, lens, lens_host, glens, parts, rank, world, dev, total_tokens,
                                    barrier)

    # ---- BASELINE configs 3 / 5 (one GPU) and 4 (any N): per-op records, driver-visible ------------------
    del h_data, h_back, h_sum, h_max, d_in
    torch.cuda.empty_cache()
    configs = {}
    if not args.no_configs:
        from benchmarks import cfg4 as cfg4_bench
        configs['cfg4'] = cfg4_bench.run(rank, world, dev, micro_batches=args.cfg4_micro_batches)
        if world == 1:
            from benchmarks import ops as ops_bench
            for key, fn in (('cfg2', lambda r, **kw: ops_bench.cfg2(r, light=True, **kw)), ('cfg3', ops_bench.cfg3),
                            ('cfg5', ops_bench.cfg5)):
                rows = []
                fn(rows, reps=5, quiet=True)
                configs[key] = ops_bench.compact(rows)
                torch.cuda.empty_cache()

    # ---- the reference on the SAME GPU and on the host cores (rank 0, single-GPU run only) -----------------
    cpu = ref_cuda = port = None
    if rank == 0 and world == 1 and not args.no_cpu:
        torch.cuda.synchronize()
        r = reference_subprocess(['--ref-device', 'cuda', '--steps', '5', '--warmup', '5'], timeout_s=300)
        if 'unavailable' in r:
            ref_cuda = r
        else:
            ref_cuda = {'value': r['value'], 'unit': 'tokens/s', 'ms_per_step': r['ms_per_step'],
                        'sample': r['cpu_baseline']['sample'], 'steps': 5, 'warmup': 5,
                        'what': 'the UNMODIFIED reference (oracle/_ref) running the same step on the same B200 through '
                                'its stock ATen CUDA path, CUDA events, separate process',
                        'speedup_of_value': value / r['value'],
                        'per_op_ms': r.get('per_op_ms'),
                        'ours_per_launch_ms': {'row_map (mean of C->P, P->L, L->R, R->C)': rm[0] / max(rm[2], 1),
                                               'segment_reduce (mean of sum, max)': (sr[0] / sr[2]) if sr else None}}
        r = reference_subprocess(['--steps', '3', '--warmup', '1', '--ref-seqs', '512'], timeout_s=300)
        if 'unavailable' in r:
            cpu = None
        else:
            cpu = dict(r['cpu_baseline'])
            cpu['sample'] += ', 3 steps'
        tps, cms, n_seq, cn, threads = time_cpu(steps=3, warmup=1, budget_s=25.0)
        port = {'value': tps, 'unit': 'tokens/s', 'cores': threads, 'kind': 'port', 'ms_per_step': cms,
                'sample': sample_text(n_seq, cn) + ', 3 steps',
                'what': 'oracle/rua_oracle.c (C + OpenMP restatement of the reference path, all host threads)'}
        if cpu is None:
            cpu = port

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': 'tokens/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms_total / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': make_config(world),
            'exchange': args.exchange if world > 1 else 'none', 'tokens_per_gpu': n_tok,
            'roofline': roofline, 'kernels': kernels, 'cpu_baseline': cpu, 'cpu_port': port,
            'reference_cuda': ref_cuda, 'configs': configs,
            'e2e': {'value': e2e_value, 'unit': 'tokens/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'steps': e2e_steps, 'ms_per_step': float(ems) / e2e_steps,
                    'what': 'every step: pinned host C batch -> H2D -> same step through the public API -> D2H of the '
                            'round-tripped C data, segment_sum and segment_max; copy-in / compute / copy-out on three '
                            'streams so that neighbouring steps overlap (PCIe-bound: 2.1 GB each way per step)'},
            'gpu_launches': int(launches), 'clocks': clk, 'host_affinity': numa,
        }
        if gather is not None:
            line['output_gather'] = gather
            # the curve WITH the gather SURVEY.md 8e names, beside `value` (which stays the no-gather, weak-scaling number)
            line['value_with_output_gather'] = {'value': gather['fused_peer_store']['value'], 'unit': 'tokens/s',
                                                'ms_per_step': gather['fused_peer_store']['ms_per_step'],
                                                'bound': 'NVLink ingress: (N-1)/N of the global payload per rank per step'}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def time_output_gather(args, step, data, lens, lens_host, glens, parts, rank, world, dev, total_tokens, barrier):
    """tokens/s WITH the final gather of outputs (every rank receives the global C data in
    original sequence order plus the global segment_sum / segment_max), two ways:
      fused   the last conversion of the step (R -> C) stores each row ONCE into the windows of all ranks over
              NVLink peer mappings (rua_row_map_multi) and into the local C copy the reductions read; the
              reductions' rows follow through rua_scatter_rows_multi.  Two 4-byte all-reduces as fences.
      nccl    step(), then all_gather of padded shards + index_put permutation (shard.gather_catted /
              gather_rows_by_sequence): what one gets from NCCL + torch ops alone.
    `value` of the bench line stays the no-gather number (the gather is wire-bound: (G-1)/G * N * D bytes per
    rank over NVLink, not HBM)."""
    import torch
    import torch.distributed as dist

    import torchrua_b200 as rua
    from torchrua_b200 import _native, shard
    steps = max(3, min(args.steps, 8))
    b_total = int(glens.numel())
    row = HIDDEN * 2
    sum_off = (total_tokens * row + 255) // 256 * 256
    max_off = sum_off + (b_total * row + 255) // 256 * 256
    windows = shard.PeerWindows(max_off + b_total * row)
    glens_dev = glens.to(dev)

    # micro-batches of the local shard: consecutive runs of sequences with ~equal token counts.  The fused R -> C +
    # peer-store kernel of micro-batch k runs on a side stream (NVLink-bound) while the conversions of micro-batch k+1
    # and the reductions of micro-batch k-1 run on the main stream (HBM-bound): the wire never waits for compute.
    cuts = shard.micro_batch_cuts(lens_host, args.gather_micro_batches)
    n_local = int(lens_host.sum())
    ids_local = parts[rank].to(dev)
    side = torch.cuda.Stream(dev, priority=-1)        # the wire-bound gather gets CTA slots as soon as they free up

    def fused_step():
        _native._CACHE.clear()
        cur = torch.cuda.current_stream()
        windows.fence()                                   # peers have finished reading last step's window
        goff, _ = _native.scan(glens_dev)                 # where every sequence of the GLOBAL batch starts
        back = torch.empty((n_local, HIDDEN), dtype=data.dtype, device=dev)
        keep = []
        for k, (a, b, ta, tb) in enumerate(cuts):
            right = rua.C(data=data[ta:tb], token_sizes=lens[a:b]).pack().left(0).right(0)
            ready = torch.cuda.Event()
            ready.record(cur)
            with torch.cuda.stream(side):
                side.wait_event(ready)
                shard.gather_catted_fused(right, parts, glens_dev, windows, fence=False, seq_ids=ids_local[a:b],
                                          goff=(goff, total_tokens), local_out=back[ta:tb])
            keep.append(right)                            # allocated on `cur`, read on `side`: freed after the join below
        cur.wait_stream(side)
        # the reductions run on the WHOLE local copy, exactly like the single-GPU step: bit-identical results (reducing
        # micro-batch slices would move the chunk boundaries of the fp32 accumulation)
        s = rua.segment_sum(back, lens)
        m = rua.segment_max(back, lens)
        full = windows.view((total_tokens, HIDDEN), data.dtype, 0)
        gs = shard.gather_rows_fused(s, parts, windows, offset_bytes=sum_off, fence=False)
        gm = shard.gather_rows_fused(m, parts, windows, offset_bytes=max_off, fence=False)
        windows.fence()
        del keep
        return full, gs, gm

    def nccl_step():
        back, s, m = step(data, lens)
        full = shard.gather_catted(back.data, lens, parts, glens_dev)
        gs = shard.gather_rows_by_sequence(s, parts)
        gm = shard.gather_rows_by_sequence(m, parts)
        return full, gs, gm

    def timed(fn):
        for _ in range(2):
            out = fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            out = fn()
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], dtype=torch.double, device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return out, float(ms) / steps

    (f_full, f_s, f_m), f_ms = timed(fused_step)
    if args.gather_only:
        windows.close()
        return {'fused_ms_per_step': f_ms, 'micro_batches': len(cuts), 'nvlink_out_GBs_per_rank': (world - 1) * int(lens_host.sum()) * row / (f_ms * 1e-3) / 1e9}
    (n_full, n_s, n_m), n_ms = timed(nccl_step)
    same = bool(torch.equal(f_full, n_full) and torch.equal(f_s, n_s) and torch.equal(f_m, n_m))
    # the global C data restricted to this rank's sequences is this rank's input (round trip = identity)
    del n_full, n_s, n_m
    wire = (world - 1) * int(lens_host.sum()) * row
    res = {'what': 'the same step, but every rank ends with the global C data + global segment_sum / segment_max',
           'fused_peer_store': {'ms_per_step': f_ms, 'value': total_tokens / (f_ms * 1e-3), 'unit': 'tokens/s',
                                'nvlink_out_GBs_per_rank': wire / (f_ms * 1e-3) / 1e9,
                                'micro_batches': len(cuts),
                                'ingress_floor_ms_at_900GBs': wire / 900e9 * 1e3,
                                'how': 'per micro-batch: conversions on the main stream, fused R->C + NVLink peer stores on a '
                                       'high-priority side stream with a capped grid (4 CTAs / SM), so the HBM-bound conversions of '
                                       'the next micro-batch share the SMs with the wire-bound stores'},
           'nccl_all_gather_then_permute': {'ms_per_step': n_ms, 'value': total_tokens / (n_ms * 1e-3),
                                            'unit': 'tokens/s'},
           'results_identical': same, 'steps': steps,
           'wire_bytes_out_per_rank_per_step': wire}
    windows.close()
    return res


_JSON_FD = None


def protect_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner, for
    one), so fd 1 is pointed at stderr for the whole run and the JSON line goes to the saved descriptor."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + '\n').encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline / reference_cuda legs')
    ap.add_argument('--no-configs', action='store_true', help='skip the configs[2..4] records')
    ap.add_argument('--no-port', action='store_true', help='reference arm: skip the C port beside the reference')
    ap.add_argument('--cfg4-micro-batches', type=int, default=0,
                    help='configs[3] record: micro-batches per rank to time (0 = the whole 64 M-token job)')
    ap.add_argument('--ref-device', default='cpu', choices=['cpu', 'cuda'],
                    help='reference arm: host cores (the contract arm) or the same GPU (the reference_cuda record)')
    ap.add_argument('--ref-seqs', type=int, default=0, help='reference arm: sequences per step (0 = full batch if it fits the budget)')
    ap.add_argument('--ref-budget', type=float, default=200.0, help='reference arm: seconds for the whole run')
    ap.add_argument('--no-gather', action='store_true', help='skip the output-gather legs of multi-GPU runs')
    ap.add_argument('--gather-only', action='store_true', help='diagnostic: time only the fused output-gather leg (N > 1)')
    ap.add_argument('--gather-micro-batches', type=int, default=4,
                    help='output gather: micro-batches per step (peer stores of one overlap the conversions of the next)')
    ap.add_argument('--exchange', default='peer', choices=['peer', 'nccl', 'none'],
                    help='per-step lengths exchange of multi-GPU runs: peer-window stores (default), NCCL all-gather, or none (diagnostic)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
