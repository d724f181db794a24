"""TEST / MEASUREMENT INFRASTRUCTURE -- stages the UNMODIFIED reference (speedcell4/torchrua 0.5.1, pure Python)
under oracle/_ref/ so that it travels to the GPU box (oracle/_ref/ is git-ignored, not gpurun-ignored).

    python oracle/make_ref.py            # authoring container only: needs /root/reference

Recipe: `pip install --no-index --no-deps --no-build-isolation --target oracle/_ref <copy of /root/reference>`
(the install builds a wheel, so it runs from a scratch copy under /tmp: /root/reference stays untouched), then a
MANIFEST.json with the sha256 of every installed module next to the sha256 of the file it came from.  If pip is
unavailable the package directory is copied verbatim instead (same bytes, same manifest).

Who may use oracle/_ref: tests/ (live parity oracle, in a SUBPROCESS -- oracle/ref_worker.py), bench.py's
`--impl reference` arm and its `reference_cuda` record.  Nothing under torchrua_b200/ ever imports it.
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')
REF_PKG = os.path.join(REF_DIR, 'torchrua')
MANIFEST = os.path.join(REF_DIR, 'MANIFEST.json')
SOURCE = os.environ.get('RUA_REFERENCE', '/root/reference')


def _sha(path: str) -> str:
    with open(path, 'rb') as fp:
        return hashlib.sha256(fp.read()).hexdigest()


def _walk(pkg: str):
    for base, _, files in sorted(os.walk(pkg)):
        for name in sorted(files):
            if name.endswith('.py'):
                full = os.path.join(base, name)
                yield os.path.relpath(full, pkg), full


def available() -> bool:
    return os.path.exists(os.path.join(REF_PKG, '__init__.py'))


def stage(force: bool = False):
    """-> path of oracle/_ref (None when the reference is neither staged nor mounted)."""
    src_pkg = os.path.join(SOURCE, 'torchrua')
    if available() and not force:
        return REF_DIR
    if not os.path.isdir(src_pkg):
        return REF_DIR if available() else None
    shutil.rmtree(REF_DIR, ignore_errors=True)
    os.makedirs(REF_DIR, exist_ok=True)
    how = 'pip install --no-index --no-deps --no-build-isolation --target oracle/_ref'
    with tempfile.TemporaryDirectory(prefix='rua_ref_src_') as tmp:
        work = os.path.join(tmp, 'src')
        shutil.copytree(SOURCE, work, ignore=shutil.ignore_patterns('.git', '__pycache__', '.hypothesis'))
        rc = subprocess.call([sys.executable, '-m', 'pip', 'install', '--quiet', '--no-index', '--no-deps',
                              '--no-build-isolation', '--find-links', '/opt/wheelhouse', '--target', REF_DIR, work],
                             stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    if rc != 0 or not available():
        how = 'verbatim copy of the package directory (pip install failed)'
        shutil.rmtree(REF_PKG, ignore_errors=True)
        shutil.copytree(src_pkg, REF_PKG, ignore=shutil.ignore_patterns('__pycache__'))
    files = {}
    for rel, full in _walk(REF_PKG):
        origin = os.path.join(src_pkg, rel)
        files[rel] = {'sha256': _sha(full), 'source_sha256': _sha(origin) if os.path.exists(origin) else None}
    unmodified = all(v['sha256'] == v['source_sha256'] for v in files.values())
    with open(MANIFEST, 'w') as fp:
        json.dump({'package': 'torchrua', 'version': '0.5.1', 'source': SOURCE, 'how': how,
                   'unmodified': unmodified, 'files': files}, fp, indent=1)
    if not unmodified:
        raise RuntimeError('oracle/_ref differs from the reference sources')
    return REF_DIR


def verify() -> bool:
    """the staged copy still has the bytes recorded at staging time (runs anywhere, no /root/reference needed)."""
    if not (available() and os.path.exists(MANIFEST)):
        return False
    files = json.load(open(MANIFEST))['files']
    seen = dict(_walk(REF_PKG))
    return set(seen) == set(files) and all(_sha(seen[rel]) == files[rel]['sha256'] for rel in files)


if __name__ == '__main__':
    path = stage(force=True)
    print(path, 'verified' if verify() else 'NOT VERIFIED')
