"""Builds libparrot_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libparrot_b200.so')
SOURCES = ['api.cu']
HEADERS = ['ptx.cuh', 'engine.cuh', 'kernels.cuh', os.path.join('..', '..', 'include', 'parrot_b200.h')]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in SOURCES + HEADERS:
        p = os.path.join(CSRC, f)
        if os.path.exists(p) and os.path.getmtime(p) > t:
            return True
    return False


def build(force=False, verbose=False):
    """Compile every CUDA source of the package for sm_100a.  Returns the library path.
    Several processes may get here at once (torchrun ranks importing a stale tree): the build runs under an exclusive
    file lock, the library is written to a temporary name and renamed, and a rank that waited re-checks staleness."""
    if not force and not _stale():
        return LIB
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('parrot_b200: nvcc not found and %s is missing or stale' % LIB)
    import fcntl
    with open(LIB + '.lock', 'w') as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():
                return LIB
            tmp = LIB + '.tmp.%d' % os.getpid()
            cmd = [nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
                   '-Xcompiler', '-fPIC', '-shared', '-o', tmp] + [os.path.join(CSRC, s) for s in SOURCES] + \
                  ['-lcuda', '-ldl']
            if verbose:
                print(' '.join(cmd))
            subprocess.check_call(cmd)
            os.replace(tmp, LIB)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB
