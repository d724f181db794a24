"""Client of oracle/ref_worker.py: the unmodified reference, live, in a subprocess (test infrastructure)."""
import os
import pickle
import struct
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, 'oracle', 'ref_worker.py')
SCENARIOS = os.path.join(ROOT, 'tests', 'scenarios.py')
REF_PKG = os.path.join(ROOT, 'oracle', '_ref', 'torchrua', '__init__.py')


def reference_staged() -> bool:
    return os.path.exists(REF_PKG)


class RefWorker:
    def __init__(self):
        env = dict(os.environ)
        env.pop('PYTHONPATH', None)
        env['PYTHONDONTWRITEBYTECODE'] = '1'
        self.proc = subprocess.Popen([sys.executable, WORKER, SCENARIOS], stdin=subprocess.PIPE,
                                     stdout=subprocess.PIPE, env=env, cwd=os.path.join(ROOT, 'oracle'))
        self.hello = self._recv()

    def _recv(self):
        head = self.proc.stdout.read(8)
        if len(head) < 8:
            raise RuntimeError(f'reference worker died (exit code {self.proc.poll()})')
        (n,) = struct.unpack('<Q', head)
        msg = pickle.loads(self.proc.stdout.read(n))
        if not msg['ok']:
            raise RuntimeError('reference worker raised:\n' + msg['error'])
        return msg['out']

    def call(self, scenario: str, device: str = 'cpu', **kwargs):
        blob = pickle.dumps({'scenario': scenario, 'device': device, 'kwargs': kwargs})
        self.proc.stdin.write(struct.pack('<Q', len(blob)))
        self.proc.stdin.write(blob)
        self.proc.stdin.flush()
        return self._recv()

    def close(self):
        if self.proc.poll() is None:
            try:
                self.proc.stdin.close()
                self.proc.wait(timeout=10)
            except Exception:
                self.proc.kill()
