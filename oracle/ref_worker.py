"""TEST / MEASUREMENT INFRASTRUCTURE -- serves the UNMODIFIED reference (oracle/_ref/torchrua, staged by
oracle/make_ref.py) from a separate process.

Why a process: importing the reference patches ``torch.Tensor.__getitem__`` / ``__setitem__`` and the shared
``torch.nn.utils.rnn.PackedSequence`` class process-wide (torchrua/core/get.py:18, layout/pack.py), and so does the
package under test; the two cannot live in one interpreter.

Protocol (binary pipes): 8-byte little-endian length + pickle.  Request ``{'scenario': name, 'device': 'cuda'|'cpu',
'kwargs': {...}}``; the scenario is looked up in the module given on the command line (tests/scenarios.py) and called
as ``fn(rua, torch.device(device), **kwargs)`` with ``rua`` = the reference.  The reply is ``{'ok': True, 'out': ...}``
(CPU tensors / plain Python) or ``{'ok': False, 'error': traceback}``.

    python oracle/ref_worker.py tests/scenarios.py
"""
import importlib.util
import os
import pickle
import struct
import sys
import traceback

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')


def import_reference():
    """the reference, and only the reference: the repo root (which holds the drop-in alias package of the same
    name) must not shadow it."""
    root = os.path.dirname(HERE)
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or os.getcwd()) != root]
    sys.path.insert(0, REF_DIR)
    import torchrua
    where = os.path.dirname(os.path.abspath(torchrua.__file__))
    if os.path.dirname(where) != REF_DIR:
        raise ImportError(f'expected the reference from {REF_DIR}, got {where}')
    return torchrua


def load_module(path: str):
    spec = importlib.util.spec_from_file_location('rua_scenarios', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _read(fp):
    head = fp.read(8)
    if len(head) < 8:
        return None
    (n,) = struct.unpack('<Q', head)
    return pickle.loads(fp.read(n))


def _write(fp, obj):
    blob = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
    fp.write(struct.pack('<Q', len(blob)))
    fp.write(blob)
    fp.flush()


def main():
    sys.dont_write_bytecode = True
    inp, out = sys.stdin.buffer, os.fdopen(os.dup(1), 'wb')
    os.dup2(2, 1)                       # anything a library prints goes to stderr, not into the pipe
    import torch
    rua = import_reference()
    scenarios = load_module(sys.argv[1])
    _write(out, {'ok': True, 'out': {'torch': torch.__version__, 'cuda': torch.cuda.is_available(),
                                     'reference': os.path.dirname(rua.__file__)}})
    while True:
        req = _read(inp)
        if req is None or req.get('scenario') == '__exit__':
            return
        try:
            fn = getattr(scenarios, req['scenario'])
            dev = torch.device(req.get('device', 'cpu'))
            res = fn(rua, dev, **req.get('kwargs', {}))
            if dev.type == 'cuda':
                torch.cuda.synchronize()
            _write(out, {'ok': True, 'out': res})
        except Exception:
            _write(out, {'ok': False, 'error': traceback.format_exc()})


if __name__ == '__main__':
    main()
