"""Builds libiiseg.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m iterative_inference_segm_b200.csrc.build [--force] [--verbose]
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ['abi.cu', 'conv_tcgen05.cu', 'pool.cu', 'update.cu', 'metrics.cu', 'layout.cu', 'deconv.cu', 'densenet.cu', 'train.cu']
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


def build(force=False, verbose=False):
    """Compile every .cu into one shared library; skipped when sources are unchanged."""
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == dig:
                return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc] + NVCC_FLAGS + ['-o', LIB] + [os.path.join(HERE, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed (exit %d)' % res.returncode)
    with open(STAMP, 'w') as fh:
        fh.write(dig)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
