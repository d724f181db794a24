"""Shared helpers for the parity tests: golden loading, digests, oracle <-> torch bridges."""
import hashlib
import os

import numpy as np

from oracle import rua_oracle as ora

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def digest(a: np.ndarray) -> np.ndarray:
    """Must match tests/golden/make_golden.py:digest."""
    h = hashlib.sha256()
    h.update(str(a.dtype).encode())
    h.update(str(tuple(a.shape)).encode())
    h.update(np.ascontiguousarray(a).tobytes())
    return np.frombuffer(h.digest(), dtype=np.uint8).copy()


class Golden:
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN, name + '.npz'))
        self.keys = set(self.z.files)

    def has(self, key):
        return key in self.keys or ('sha:' + key) in self.keys

    def names(self, prefix=''):
        return sorted({k[4:] if k.startswith('sha:') else k for k in self.keys if
                       (k[4:] if k.startswith('sha:') else k).startswith(prefix)})

    def __getitem__(self, key):
        return self.z[key]

    def check(self, key, actual: np.ndarray, exact=True, rtol=0.0, atol=0.0):
        actual = np.asarray(actual)
        if key in self.keys:
            expected = self.z[key]
            assert actual.shape == expected.shape, f'{self.name}:{key} shape {actual.shape} != {expected.shape}'
            assert actual.dtype == expected.dtype, f'{self.name}:{key} dtype {actual.dtype} != {expected.dtype}'
            if exact:
                same = (actual == expected) | (_isnan(actual) & _isnan(expected))
                assert same.all(), f'{self.name}:{key} differs in {int((~same).sum())} of {same.size} entries'
            else:
                np.testing.assert_allclose(actual, expected, rtol=rtol, atol=atol, equal_nan=True,
                                           err_msg=f'{self.name}:{key}')
        elif 'sha:' + key in self.keys:
            assert exact, 'digest-only golden entries are bit-exact by construction'
            assert (digest(actual) == self.z['sha:' + key]).all(), f'{self.name}:{key} digest mismatch'
        else:
            raise KeyError(f'{self.name}:{key}')

    def seq(self, prefix, kind):
        """oracle container for a sequence stored with Rec.put_seq (full data only)."""
        data = self.z[prefix + '.data']
        if kind == 'P':
            return ora.Pack(data, self.z[prefix + '.batch_sizes'], self.z[prefix + '.sorted_indices'],
                            self.z[prefix + '.unsorted_indices'])
        cls = {'C': ora.Cat, 'L': ora.Left, 'R': ora.Right}[kind]
        return cls(data, self.z[prefix + '.token_sizes'])


def _isnan(a):
    return np.isnan(a) if np.issubdtype(a.dtype, np.floating) else np.zeros(a.shape, dtype=bool)
