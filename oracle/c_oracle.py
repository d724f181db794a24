"""ctypes wrapper over oracle/rua_oracle.c (built into oracle/_build/librua_oracle.so by
__graft_entry__.build()).  TEST INFRASTRUCTURE ONLY -- the multi-threaded CPU restatement used for
full-size checks and as bench.py's CPU baseline.  The product package never imports this."""
import ctypes
import os
import subprocess
from ctypes import c_int, c_int64, c_void_p

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'rua_oracle.c')
SO = os.path.join(HERE, '_build', 'librua_oracle.so')

CAT, LEFT, PACK, RIGHT = 0, 1, 2, 3
KIND = {'C': CAT, 'L': LEFT, 'P': PACK, 'R': RIGHT}
OPS = {'sum': 0, 'mean': 1, 'prod': 2, 'max': 3, 'min': 4, 'logsumexp': 5}
MAPS = {'id': 0, 'rev': 1, 'roll': 2}

_lib = None


def build(force=False):
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(['gcc', '-O3', '-march=native', '-fopenmp', '-fPIC', '-shared', '-std=c11', SRC, '-o', SO,
                               '-lm'])
    return SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            build()
        _lib = ctypes.CDLL(SO)
        _lib.ora_num_threads.restype = c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(c_void_p)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def num_threads() -> int:
    return int(lib().ora_num_threads())


def use_all_cores() -> int:
    """torchrun exports OMP_NUM_THREADS=1; the CPU baseline is meant to use every core it may run on."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().ora_set_num_threads(c_int(n))
    return num_threads()


def pack_meta(lens, sorted_in=None):
    lens = _i64(lens)
    B = lens.shape[0]
    T = int(lens.max()) if B else 0
    bs = np.empty(T, dtype=np.int64)
    srt = np.empty(B, dtype=np.int64)
    uns = np.empty(B, dtype=np.int64)
    sin = None if sorted_in is None else _i64(sorted_in)
    lib().ora_pack_meta(_p(lens), c_int64(B), c_int64(T), _p(sin), _p(bs), _p(srt), _p(uns))
    return bs, srt, uns


def lengths_from_pack(bs, unsorted):
    bs, unsorted = _i64(bs), _i64(unsorted)
    out = np.empty(unsorted.shape[0], dtype=np.int64)
    lib().ora_lengths_from_pack(_p(bs), _p(unsorted), c_int64(unsorted.shape[0]), c_int64(bs.shape[0]), _p(out))
    return out


def move(src: np.ndarray, src_kind: str, dst_kind: str, lens, bs=None, unsorted=None, fill=0, mapping='id',
         shift=0, pad_row0=False, out=None):
    """src: flattened-or-not storage of the source layout; returns the destination storage
    ((N,*) for C/P, (B,T,*) for L/R).  The fill value is given in the payload dtype."""
    lens = _i64(lens)
    B = lens.shape[0]
    T = int(lens.max()) if B else 0
    N = int(lens.sum())
    feat = src.shape[2:] if src_kind in 'LR' else src.shape[1:]
    row_bytes = int(np.prod(feat, dtype=np.int64)) * src.dtype.itemsize
    src_w = src.shape[1] if src_kind in 'LR' else 0
    shape = (B, T) + tuple(feat) if dst_kind in 'LR' else (N,) + tuple(feat)
    if out is None:
        out = np.empty(shape, dtype=src.dtype)
    fillv = np.asarray([fill], dtype=src.dtype)
    src = np.ascontiguousarray(src)
    bs = None if bs is None else _i64(bs)
    unsorted = None if unsorted is None else _i64(unsorted)
    Tp = 0 if bs is None else bs.shape[0]
    lib().ora_move(_p(src), _p(out), c_int64(row_bytes), c_int(KIND[src_kind]), c_int(KIND[dst_kind]), _p(lens),
                   c_int64(B), _p(bs), c_int64(Tp), _p(unsorted), c_int64(src_w), c_int64(T), c_int(MAPS[mapping]),
                   c_int64(shift), _p(fillv), c_int(src.dtype.itemsize), c_int(int(pad_row0)))
    return out


def mask(lens, width, zero, one, dtype):
    lens = _i64(lens)
    out = np.empty((lens.shape[0], width), dtype=dtype)
    z, o = np.asarray([zero], dtype=dtype), np.asarray([one], dtype=dtype)
    lib().ora_mask(_p(lens), c_int64(lens.shape[0]), c_int64(width), _p(z), _p(o), c_int(out.dtype.itemsize), _p(out))
    return out


def segment_reduce(data: np.ndarray, sizes, op: str, bf16=False):
    """data: float32 (N, H...) or, with bf16=True, uint16 bit patterns; returns the same kind."""
    sizes = _i64(sizes)
    S = sizes.shape[0]
    data = np.ascontiguousarray(data)
    assert data.dtype == (np.uint16 if bf16 else np.float32)
    H = int(np.prod(data.shape[1:], dtype=np.int64))
    out = np.empty((S,) + data.shape[1:], dtype=data.dtype)
    lib().ora_segment_reduce(_p(data), _p(sizes), c_int64(S), c_int64(H), c_int(OPS[op]), c_int(int(bf16)), _p(out))
    return out
