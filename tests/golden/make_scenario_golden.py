"""Freeze the outputs of the UNMODIFIED reference on the scenario cases of tests/scenario_cases.py.

Run in the authoring container only:

    cd /tmp && PYTHONDONTWRITEBYTECODE=1 python /root/repo/tests/golden/make_scenario_golden.py

Every array in tests/golden/scenarios.npz is an output of speedcell4/torchrua 0.5.1 (imported from /root/reference or,
failing that, from the staged copy oracle/_ref) on CPU, torch 2.11.0+cu128, produced by the scenario functions of
tests/scenarios.py -- the same functions the GPU tests run on the package under test.  Keys: '<case>|<path><tag>'.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
ROOT = os.path.dirname(TESTS)
REF = os.environ.get('RUA_REFERENCE', '/root/reference')
if not os.path.isdir(os.path.join(REF, 'torchrua')):
    REF = os.path.join(ROOT, 'oracle', '_ref')
sys.path[:] = [p for p in sys.path if os.path.abspath(p or os.getcwd()) != ROOT]
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

import torch  # noqa: E402
import torchrua as rua  # noqa: E402  (the reference)

assert os.path.dirname(os.path.dirname(os.path.abspath(rua.__file__))) == os.path.abspath(REF), rua.__file__
sys.path.insert(1, TESTS)
import scenarios  # noqa: E402
from scenario_cases import CASES  # noqa: E402
from treecmp import flatten, freeze  # noqa: E402


def main():
    torch.set_num_threads(1)
    rec = {}
    for case, fn, kwargs, mode, ref_kwargs in CASES:
        kw = dict(kwargs)
        kw.update(ref_kwargs or {})
        tree = getattr(scenarios, fn)(rua, torch.device('cpu'), **kw)
        n = 0
        for path, leaf in flatten(tree).items():
            arr, tag = freeze(leaf)
            rec[f'{case}|{path}{tag}'] = arr
            n += 1
        print(f'{case}: {n} leaves')
    out = os.path.join(HERE, 'scenarios.npz')
    np.savez_compressed(out, **rec)
    print(f'{len(rec)} arrays, {os.path.getsize(out) / 1024:.1f} KiB -> {out}')


if __name__ == '__main__':
    main()
