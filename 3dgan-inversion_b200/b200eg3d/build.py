"""Build libb200eg3d.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libb200eg3d.so')
SOURCES = ['bank.cu', 'conv_tc.cu', 'conv_api.cu', 'conv_simt.cu', 'modconv.cu', 'elementwise.cu', 'triplane.cu', 'triplane_tc.cu', 'raymarch.cu', 'losses.cu', 'projector.cu', 'optim.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC,-fvisibility=hidden']


def _nvcc():
    for c in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if c and os.path.exists(c):
            return c
    raise RuntimeError('nvcc not found')


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in os.listdir(CSRC))


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    flags = list(NVCC_FLAGS)
    procs = []
    objs = []
    for s in SOURCES:
        o = os.path.join(objdir, s.replace('.cu', '.o'))
        objs.append(o)
        src = os.path.join(CSRC, s)
        if not force and os.path.exists(o) and os.path.getmtime(o) > max(
                os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC) if f.endswith(('.h', '.cuh')) or f == s):
            continue
        cmd = [nvcc] + flags + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f'nvcc failed on {s}')
    cmd = [nvcc, '-shared', '-o', LIB] + objs + ['-lcudart']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout)
        raise RuntimeError('link failed')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
