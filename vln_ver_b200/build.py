"""Builds vln_ver_b200/libver_b200.so in-tree with nvcc for sm_100a (cross-compiles
without a GPU).  `python -m vln_ver_b200.build [--force] [--verbose]`."""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
INCLUDE = os.path.join(os.path.dirname(HERE), 'include')
LIB = os.path.join(HERE, 'libver_b200.so')
STAMP = os.path.join(HERE, '.libver_b200.stamp')

SOURCES = ['api.cu', 'geometry.cu', 'msda.cu', 'msda3d.cu', 'col2im.cu', 'sca.cu', 'sca_tc.cu', 'sca_tc3.cu', 'sca_tc4.cu', 'sca_tc5.cu', 'sca_tc6.cu', 'sca_tc7.cu', 'sca_bwd_tc2.cu', 'order.cu', 'gemm_tc.cu', 'elementwise.cu',
           'fused_norm.cu']
NVCC_FLAGS = ['--threads', '8', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '--shared']


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, INCLUDE):
        for name in sorted(os.listdir(root)):
            if name.endswith(('.cu', '.cuh', '.h')):
                with open(os.path.join(root, name), 'rb') as f:
                    h.update(name.encode() + b'\0' + f.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return 'nvcc'


def build(force=False, verbose=False):
    """Compile every .cu into one shared library; no-op when sources are unchanged."""
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as f:
            if f.read().strip() == digest:
                return LIB
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = [nvcc_path()] + NVCC_FLAGS + ['-I', INCLUDE] + (['-Xptxas', '-v'] if verbose else []) + ['-o', LIB] + srcs
    if verbose:
        print(' '.join(cmd))
    subprocess.run(cmd, check=True)
    with open(STAMP, 'w') as f:
        f.write(digest)
    return LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv, verbose='--verbose' in sys.argv)
    print(LIB)
