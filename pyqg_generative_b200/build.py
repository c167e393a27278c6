"""Build libqgb200.so in-tree with nvcc for sm_100a:  python -m pyqg_generative_b200.build

Every ``csrc/*.cu`` file is one translation unit (compiled in parallel, ``--split-compile`` inside each); the objects are kept
under ``build/`` and only stale ones are rebuilt.  Staleness is decided by CONTENT hashes (source + headers + flags), not
by mtimes: the tree is snapshotted onto the GPU box by tools that do not preserve timestamps, and a spurious rebuild there
would burn minutes of GPU time.
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'libqgb200.so')
STAMP = OUT + '.hash'
OBJDIR = os.path.join(os.path.dirname(HERE), 'build', 'obj')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC']


def _sha(paths, extra=''):
    h = hashlib.sha256(extra.encode())
    for p in sorted(paths):
        h.update(os.path.basename(p).encode())
        with open(p, 'rb') as f:
            h.update(f.read())
    return h.hexdigest()


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.hpp', '.h'))]
    hs.append(os.path.join(os.path.dirname(HERE), 'include', 'qgb200.h'))
    return sorted(hs)


def _deps(path, seen=None):
    """``path`` plus the in-tree headers it includes, recursively (quoted includes only)."""
    import re
    seen = set() if seen is None else seen
    path = os.path.normpath(path)
    if path in seen or not os.path.exists(path):
        return seen
    seen.add(path)
    with open(path) as f:
        for inc in re.findall(r'^\s*#\s*include\s+"([^"]+)"', f.read(), flags=re.M):
            _deps(os.path.join(os.path.dirname(path), inc), seen)
    return seen


def _tu_hash(src):
    return _sha(_deps(src), ' '.join(FLAGS))


def tree_hash():
    return _sha(_sources() + _headers(), ' '.join(FLAGS))


def needs_build():
    if not os.path.exists(OUT) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != tree_hash()


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    os.makedirs(OBJDIR, exist_ok=True)
    meta_path = os.path.join(OBJDIR, 'hashes.json')
    try:
        with open(meta_path) as f:
            meta = json.load(f)
    except Exception:
        meta = {}
    jobs, objs = [], []
    for src in _sources():
        obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        hsh = _tu_hash(src)
        if force or not os.path.exists(obj) or meta.get(os.path.basename(src)) != hsh:
            jobs.append((src, obj, hsh))

    def compile_one(job):
        src, obj, hsh = job
        cmd = [nvcc] + FLAGS + ['--split-compile=0'] + (['-Xptxas', '-v'] if verbose else []) + ['-c', '-o', obj, src]
        print(' '.join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
        return os.path.basename(src), hsh

    if jobs:
        with ThreadPoolExecutor(max_workers=min(4, len(jobs))) as ex:
            for name, hsh in ex.map(compile_one, jobs):
                meta[name] = hsh
        with open(meta_path, 'w') as f:
            json.dump(meta, f)
    cmd = [nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', OUT] + objs
    print(' '.join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    with open(STAMP, 'w') as f:
        f.write(tree_hash())
    return OUT


if __name__ == '__main__':
    build(force='--force' in sys.argv, verbose='-v' in sys.argv)
