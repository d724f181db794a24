"""The bench line the driver parses: every key of the measurement contract must be present in the latest committed
B200 lines (profiles/r2_bench_*gpu.json are verbatim copies of what `python bench.py` printed on the GPU box), and
frac / e2e / launches must be self-consistent.  CPU-only: guards the schema, not the numbers."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINES = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r2_bench_*gpu.json')))

TOP = ['metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
       'dtype', 'data', 'config', 'roofline', 'e2e', 'gpu_launches', 'clocks']


@pytest.mark.parametrize('path', LINES, ids=[os.path.basename(p) for p in LINES])
def test_committed_bench_line_has_the_contract_keys(path):
    d = json.load(open(path))
    for k in TOP:
        assert k in d, k
    assert d['higher_is_better'] is True and d['scaling'] == 'weak' and d['vs_baseline'] is None
    assert 'workload' in d['config'] and 'model' not in d['config']
    r = d['roofline']
    for k in ('bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'):
        assert k in r, k
    assert r['bound'] == 'hbm' and r['unit'] == 'GB/s'
    assert abs(r['frac'] - r['achieved'] / r['peak']) < 1e-6 and 0 < r['frac'] <= 1.0
    e = d['e2e']
    for k in ('value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step'):
        assert k in e, k
    assert e['h2d_bytes_per_step'] > 0 and e['d2h_bytes_per_step'] > 0 and e['value'] < d['value']
    assert d['gpu_launches'] > 0 and d['gpu_launches'] % d['steps'] == 0
    c = d['clocks']
    for k in ('sm_mhz', 'sm_max_mhz', 'reasons'):
        assert k in c, k
    if d['n_gpus'] == 1:
        b = d['cpu_baseline']
        for k in ('value', 'unit', 'cores', 'kind', 'sample'):
            assert k in b, k
        assert b['kind'] in ('reference', 'port')
    # value = tokens of all ranks / max-over-ranks step time
    assert abs(d['value'] - d['config']['tokens_total'] / (d['ms_per_step'] * 1e-3)) / d['value'] < 1e-3


def test_there_is_a_committed_line_per_gpu_count():
    assert {os.path.basename(p) for p in LINES} >= {f'r2_bench_{n}gpu.json' for n in (1, 2, 4, 8)}
