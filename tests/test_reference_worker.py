"""CPU: the staged reference (oracle/_ref, served by oracle/ref_worker.py) is the unmodified reference, and running the
shared scenario functions on it reproduces the frozen outputs in tests/golden/scenarios.npz bit for bit -- i.e. the two
reference-side oracles the GPU tests compare against (live and frozen) are one and the same thing."""
import os

import numpy as np
import pytest

from tests.refclient import RefWorker, reference_staged
from tests.scenario_cases import CASES
from tests.treecmp import compare, thaw

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'scenarios.npz')
needs_ref = pytest.mark.skipif(not reference_staged(), reason='oracle/_ref is not staged (python oracle/make_ref.py)')


@needs_ref
def test_staged_reference_is_unmodified():
    from oracle import make_ref
    assert make_ref.verify(), 'oracle/_ref no longer matches the sha256 manifest written when it was staged'
    import json
    man = json.load(open(make_ref.MANIFEST))
    assert man['unmodified'] and man['version'] == '0.5.1' and len(man['files']) >= 20


@pytest.fixture(scope='module')
def worker():
    w = RefWorker()
    yield w
    w.close()


@pytest.fixture(scope='module')
def frozen():
    z = np.load(GOLDEN)
    trees = {}
    for key in z.files:
        case, path = key.split('|', 1)
        tag = ''
        if path.endswith('#bf16'):
            path, tag = path[:-5], '#bf16'
        trees.setdefault(case, {})[path] = thaw(z[key], tag)
    return trees


@needs_ref
@pytest.mark.parametrize('case,fn,kwargs,mode,ref_kwargs', CASES, ids=[c[0] for c in CASES])
def test_live_reference_on_cpu_reproduces_the_frozen_outputs(worker, frozen, case, fn, kwargs, mode, ref_kwargs):
    from tests.treecmp import flatten
    assert 'oracle/_ref' in worker.hello['reference'].replace(os.sep, '/')
    kw = dict(kwargs)
    kw.update(ref_kwargs or {})
    got = flatten(worker.call(fn, device='cpu', **kw))
    compare('exact', got, frozen[case], label=case)
