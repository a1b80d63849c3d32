"""Builds libwc_b200.so in-tree with nvcc for sm_100a (the only target)."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libwc_b200.so')
STAMP = os.path.join(HERE, '.libwc_b200.stamp')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Wno-deprecated-gpu-targets']

def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _digest() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), 'include')):
        for f in sorted(os.listdir(root)):
            if f.endswith(('.cu', '.cuh', '.h')):
                with open(os.path.join(root, f), 'rb') as fh:
                    h.update(f.encode())
                    h.update(fh.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path() -> str:
    for cand in (shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found: libwc_b200.so cannot be built')


def _fresh(dig: str) -> bool:
    if os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            return fh.read().strip() == dig
    return False


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ into one shared library.  Returns its path.  A no-op when the digest of the
    sources matches the stamp; concurrent callers (the ranks of one torchrun launch) are serialised with a file lock
    and the library is moved into place atomically."""
    import fcntl
    dig = _digest()
    if not force and _fresh(dig):
        return LIB
    with open(os.path.join(HERE, '.build.lock'), 'w') as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and _fresh(dig):       # another rank built it while we waited
            return LIB
        return _build_locked(dig, verbose)


def _build_locked(dig: str, verbose: bool) -> str:
    objdir = os.path.join(os.path.dirname(HERE), 'build')
    os.makedirs(objdir, exist_ok=True)
    nvcc = nvcc_path()
    procs = []
    objs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError('nvcc failed: ' + ' '.join(cmd))
    tmp = LIB + f'.tmp{os.getpid()}'
    cmd = [nvcc, '-shared', '-cudart', 'shared', '-o', tmp] + objs
    subprocess.run(cmd, check=True)
    os.replace(tmp, LIB)
    with open(STAMP, 'w') as fh:
        fh.write(dig)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
