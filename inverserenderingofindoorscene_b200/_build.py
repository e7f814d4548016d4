"""In-tree build of libirb_b200.so (hand-written sm_100a CUDA + the C ABI of include/irb.h).

``python -m inverserenderingofindoorscene_b200._build`` or ``__graft_entry__.build()``.
nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with gpurun.
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, 'csrc')
INCLUDE = os.path.join(os.path.dirname(PKG), 'include')
LIB = os.path.join(PKG, 'libirb_b200.so')
OBJ = os.path.join(PKG, 'build')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
FLAGS = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '-Xptxas', '-v', '-I', INCLUDE]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _digest():
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + ['../../include/irb.h']:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            h.update(open(p, 'rb').read())
    h.update(' '.join(ARCH + FLAGS).encode())
    return h.hexdigest()


def _compile(src):
    obj = os.path.join(OBJ, src[:-3] + '.o')
    cmd = [NVCC] + ARCH + FLAGS + ['-c', os.path.join(CSRC, src), '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(obj + '.log', 'w') as f:
        f.write(' '.join(cmd) + '\n' + log)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed for %s:\n%s' % (src, log))
    return obj


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, 'digest.txt')
    dig = _digest()
    if (not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig):
        return LIB
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(_compile, sources()))
    cmd = [NVCC] + ARCH + ['-shared', '-o', LIB] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stdout + r.stderr)
    with open(stamp, 'w') as f:
        f.write(dig)
    if verbose:
        for o in objs:
            print(open(o + '.log').read())
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
