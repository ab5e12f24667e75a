"""Builds liblbm_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'lbm_b200.cu')
DEPS = [SRC, os.path.join(HERE, 'csrc', 'lbm_kernels.cuh'), os.path.join(HERE, 'csrc', 'lbm_device.cuh'),
        os.path.join(os.path.dirname(HERE), 'include', 'lbm_b200.h')]
OUT = os.path.join(HERE, 'liblbm_b200.so')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-fmad=false',            # belt and braces: the kernels use __dadd_rn/__dmul_rn throughout
              '-shared', '-Xcompiler', '-fPIC', '-cudart', 'static']


def nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found')
    return exe


def build(force=False, verbose=False):
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    cmd = [nvcc()] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-o', OUT, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return OUT


if __name__ == '__main__':
    import sys
    print(build(force='-f' in sys.argv, verbose='-v' in sys.argv))
