#!/usr/bin/env python
"""Aggregate host<->device bandwidth when ALL ranks of a node copy at once (context for bench.py's e2e curve at N > 1:
every rank streams its batch through pinned host memory each step, so the e2e number is bounded by what the host side
-- PCIe root complexes, memory controllers, NUMA placement -- sustains for N GPUs together, not by the kernels).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P benchmarks/pcie_duplex_multi.py

Prints one JSON line on rank 0: per-rank and aggregate GB/s for H2D alone, D2H alone and both directions together."""
import json
import os

import torch
import torch.distributed as dist


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (('RANK', 0), ('WORLD_SIZE', 1), ('LOCAL_RANK', 0)))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    n = 1 << 30
    h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d_in = torch.empty(n, dtype=torch.uint8, device=dev)
    d_out = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(h2d, d2h, reps=6):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        s1.wait_event(a)
        s2.wait_event(a)
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)
        b.record()
        torch.cuda.synchronize()
        gbs = reps * n * (int(h2d) + int(d2h)) / 1e9 / (a.elapsed_time(b) * 1e-3)
        t = torch.tensor([gbs], dtype=torch.double, device=dev)
        if world > 1:
            out = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(out, t)
            return [float(x) for x in out]
        return [gbs]

    run(True, True, 1)
    res = {}
    for name, (a, b) in (('h2d_only', (True, False)), ('d2h_only', (False, True)), ('both', (True, True))):
        per = run(a, b)
        res[name] = {'per_rank_GBs': [round(x, 1) for x in per], 'aggregate_GBs': round(sum(per), 1)}
    if rank == 0:
        try:
            aff = sorted(os.sched_getaffinity(0))
            aff = f'{aff[0]}-{aff[-1]} ({len(aff)} cpus)'
        except Exception:
            aff = None
        print(json.dumps({'n_gpus': world, 'bytes_per_copy': n, 'cpu_affinity_rank0': aff, **res}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
