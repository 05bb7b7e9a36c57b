"""Builds libiiseg.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m iterative_inference_segm_b200.csrc.build [--force] [--verbose]
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ['abi.cu', 'conv_tcgen05.cu', 'pool.cu', 'update.cu', 'metrics.cu', 'layout.cu', 'deconv.cu', 'densenet.cu', 'train.cu', 'contextmod.cu']
HEADERS = ['common.cuh', os.path.join('..', '..', 'include', 'iiseg.h')]
LIB = os.path.join(HERE, 'libiiseg.so')
STAMP = os.path.join(HERE, '.libiiseg.stamp')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-Xptxas=-v',
              '-Xcompiler', '-fPIC,-O2', '-shared',
              '-cudart', 'static']


def _digest():
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(HERE, f), 'rb') as fh:
            h.update(fh.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _fresh(dig):
    if os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            return fh.read().strip() == dig
    return False


def build(force=False, verbose=False):
    """Compile every .cu into one shared library; skipped when sources are unchanged.  Safe under `torchrun`: the ranks
    of a node serialise on a file lock, the first one compiles (into a temporary file, renamed atomically), the others
    find the fresh stamp."""
    import fcntl
    dig = _digest()
    if not force and _fresh(dig):
        return LIB
    with open(os.path.join(HERE, '.libiiseg.lock'), 'w') as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and _fresh(dig):
                return LIB
            nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
            tmp = LIB + '.tmp.%d' % os.getpid()
            cmd = [nvcc] + NVCC_FLAGS + ['-o', tmp] + [os.path.join(HERE, s) for s in SOURCES]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError('nvcc failed (exit %d)' % res.returncode)
            os.replace(tmp, LIB)
            with open(STAMP, 'w') as fh:
                fh.write(dig)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
