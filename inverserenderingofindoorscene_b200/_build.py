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


def _digest(defines=()):
    h = hashlib.sha256()
    h.update(' '.join(defines).encode())
    for f in sorted(os.listdir(CSRC)) + ['../../include/irb.h']:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            h.update(open(p, 'rb').read())
    h.update(' '.join(ARCH + FLAGS).encode())
    return h.hexdigest()


def _compile(src, objdir=OBJ, defines=()):
    obj = os.path.join(objdir, src[:-3] + '.o')
    cmd = [NVCC] + ARCH + FLAGS + ['-D' + d for d in defines] + ['-c', os.path.join(CSRC, src), '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(obj + '.log', 'w') as f:
        f.write(' '.join(cmd) + '\n' + log)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed for %s:\n%s' % (src, log))
    return obj


def build(force=False, verbose=False, defines=(), variant=None):
    """default: libirb_b200.so.  ``variant='x', defines=('IRB_FOO=0',)`` builds libirb_b200_x.so with extra -D flags for
    A/B runs on the GPU box (select it with IRB_LIB=<path>, see _lib.py); variants are experiment artefacts, git-ignored."""
    lib = LIB if variant is None else os.path.join(PKG, 'libirb_b200_%s.so' % variant)
    objdir = OBJ if variant is None else os.path.join(OBJ, variant)
    os.makedirs(objdir, exist_ok=True)
    stamp = os.path.join(objdir, 'digest.txt')
    dig = _digest(defines)
    if (not force and os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read() == dig):
        return lib
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(lambda src: _compile(src, objdir, defines), sources()))
    cmd = [NVCC] + ARCH + ['-shared', '-o', lib] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stdout + r.stderr)
    with open(stamp, 'w') as f:
        f.write(dig)
    if verbose:
        for o in objs:
            print(open(o + '.log').read())
    return lib


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
