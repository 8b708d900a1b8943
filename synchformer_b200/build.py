"""Build the C-ABI shared library `synchformer_b200/lib/libsynchformer_b200.so` with nvcc for sm_100a.

The library is built IN-TREE (it is git-ignored but travels to the GPU box with the repo snapshot).
nvcc cross-compiles without a GPU.  Usage:  python -m synchformer_b200.build [--force]
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
LIB = os.path.join(LIBDIR, 'libsynchformer_b200.so')
SOURCES = ['runtime.cu', 'gemm_tcgen05.cu', 'layernorm.cu', 'attention.cu', 'attention_tc.cu', 'embed.cu', 'mel.cu', 'train.cu', 'attention_train.cu', 'attention_bwd.cu', 'optim.cu', 'contrastive.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-O2']


def _nvcc() -> str:
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return 'nvcc'


def _digest() -> str:
    h = hashlib.sha256(' '.join(NVCC_FLAGS).encode())
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(HERE, '..', 'include', 'synchformer_b200.h')]
    for f in files:
        with open(f, 'rb') as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, 'build.sha256')
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIB
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIBDIR, src.replace('.cu', '.o'))
        cmd = [_nvcc()] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            print(out)
        if p.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}')
    cmd = [_nvcc(), '-shared', '-o', LIB] + objs + ['-lcudart']
    subprocess.check_call(cmd)
    with open(stamp, 'w') as f:
        f.write(digest)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
