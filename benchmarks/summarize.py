#!/usr/bin/env python
"""Turn raw measurement files (gpurun_out/) into the tracked summaries under profiles/.

    python benchmarks/summarize.py ops      <ops.json> <out.md> [title]
    python benchmarks/summarize.py launches <ncu launch-list csv> <out.md> <first-kernel-regex> [title]
    python benchmarks/summarize.py ncu      <ncu --page raw --csv export> <out.md> [title]
"""
import csv
import json
import re
import sys


def ops(src, dst, title='per-op timings on one B200 (`python benchmarks/ops.py --reps 15`)'):
    rows = json.load(open(src))
    rows = rows['rows'] if isinstance(rows, dict) else rows
    out = [f'# {title}', '',
           'api = CUDA events around the public API call with the metadata cache cleared (scans, sorts and the host sync are '
           'inside); kernel = events right around the kernel launches of that call; GB/s use the ALGORITHMIC bytes of '
           'SURVEY.md 8d; peak = measured copy bandwidth (MEASURED_PEAKS.json).  `aten payload` = the stock ATen op the '
           'reference would end in, given a PRECOMPUTED index: a lower bound on the reference-on-GPU cost.', '',
           '| cfg | op | alg GB | api ms | api GB/s | kernel GB/s | kernel % of peak | Mtok/s (api) | aten payload ms | vs aten |',
           '|---|---|---|---|---|---|---|---|---|---|']
    for r in rows:
        k = r.get('kernel_GBs')
        out.append('| {cfg} | {op} | {alg:.3f} | {api:.3f} | {ag:.0f} | {kg} | {kf} | {mt:.0f} | {at} | {sp} |'.format(
            cfg=r['cfg'], op=r['op'], alg=r['alg_GB'], api=r['api_ms'], ag=r['api_GBs'],
            kg=f'{k:.0f}' if k else '-', kf=f"{100 * r['kernel_frac_of_measured_peak']:.1f}" if k else '-',
            mt=r['Mtok_s'], at=f"{r['aten_payload_ms']:.3f}" if 'aten_payload_ms' in r else '',
            sp=f"x{r['speedup_vs_aten_payload']:.1f}" if 'speedup_vs_aten_payload' in r else ''))
    open(dst, 'w').write('\n'.join(out) + '\n')


def _launch_rows(src):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    k, m, v, i = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
    out = []
    for r in rows[1:]:
        if r[m] == 'gpu__time_duration.sum':
            out.append((int(r[i]), r[k], float(r[v].replace(',', '')) / 1e3))   # us
    return out


def launches(src, dst, first, title='ncu launch list of one bench step'):
    rows = _launch_rows(src)
    starts = [n for n, (_, name, _) in enumerate(rows) if re.search(first, name)]
    assert len(starts) >= 2, 'need at least two steps in the capture'
    a, b = starts[-2], starts[-1]          # the last complete step
    step = rows[a:b]
    total = sum(us for _, _, us in step)
    out = [f'# {title}', '',
           'Per-launch times under ncu are cold-cache and serialised: compare SHARES with the event-timed bench line, not absolutes.',
           '', '| # | kernel | us | share of step |', '|---|---|---|---|']
    by = {}
    for n, (_, name, us) in enumerate(step):
        short = re.sub(r'^void ', '', name)[:90]
        out.append(f'| {n} | `{short}` | {us:.1f} | {100 * us / total:.1f} % |')
        key = re.sub(r'\(.*', '', short)
        by[key] = by.get(key, 0.0) + us
    out += ['', f'Sum {total:.0f} us over {len(step)} launches.  By kernel: ' +
            '; '.join(f'{k} {100 * v / total:.1f} %' for k, v in sorted(by.items(), key=lambda kv: -kv[1]))]
    open(dst, 'w').write('\n'.join(out) + '\n')


NCU_COLS = [('gpu__time_duration.sum', 'time'), ('dram__bytes_read.sum', 'dram read'), ('dram__bytes_write.sum', 'dram write'),
            ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram % of ncu peak'),
            ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
            ('smsp__issue_active.avg.pct', 'issue active %'), ('l1tex__throughput.avg.pct_of_peak_sustained_active', 'l1tex %'),
            ('launch__registers_per_thread', 'regs'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
            ('smsp__inst_executed.sum', 'warp instructions')]


def ncu(src, dst, title='ncu --set full summary'):
    rows = list(csv.reader(open(src)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [(hdr.index(c), label) for c, label in NCU_COLS if c in hdr]
    kn = hdr.index('Kernel Name')
    out = [f'# {title}', '', '| kernel | ' + ' | '.join(f'{label} ({units[i]})' if units[i] else label for i, label in cols) + ' |',
           '|---|' + '---|' * len(cols)]
    for r in data:
        vals = []
        for i, _ in cols:
            try:
                x = float(r[i].replace(',', ''))
                vals.append(f'{x:.3g}' if x < 1e6 else f'{x:.4g}')
            except ValueError:
                vals.append(r[i])
        out.append(f'| `{r[kn][:70]}` | ' + ' | '.join(vals) + ' |')
    open(dst, 'w').write('\n'.join(out) + '\n')


if __name__ == '__main__':
    cmd, args = sys.argv[1], sys.argv[2:]
    {'ops': ops, 'launches': launches, 'ncu': ncu}[cmd](*args)
