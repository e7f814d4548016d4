"""ctypes wrapper around tests/hostcheck/libhostcheck.so (TEST INFRASTRUCTURE): the per-pixel
math of the CUDA kernels instantiated on the host in float/double."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'hostcheck', 'hostcheck.cpp')
LIB = os.path.join(HERE, 'hostcheck', 'libhostcheck.so')
CSRC = os.path.join(os.path.dirname(HERE), 'inverserenderingofindoorscene_b200', 'csrc')
MAXD = 128

FP = C.POINTER(C.c_float)


class EnvTab(C.Structure):
    _fields_ = [('lx', C.c_float * MAXD), ('ly', C.c_float * MAXD), ('lz', C.c_float * MAXD),
                ('w', C.c_float * MAXD), ('D', C.c_int)]


class RenderSgArgs(C.Structure):
    _fields_ = [(n, FP) for n in ('albedo', 'normal', 'rough', 'axis', 'lamb', 'weight', 'view', 'out_diff',
                                  'out_spec', 'g_diff', 'g_spec', 'g_albedo', 'g_normal', 'g_rough', 'g_axis',
                                  'g_lamb', 'g_weight', 'lobe_uv_out', 'lobe_uv_in')] + \
               [('B', C.c_int), ('K', C.c_int), ('R', C.c_int), ('C', C.c_int), ('F0', C.c_float), ('warp', C.c_int)]


class SgEnvArgs(C.Structure):
    _fields_ = [(n, FP) for n in ('axis', 'lamb', 'weight', 'env', 'lamb_w', 'weight_w', 'g_env', 'g_axis',
                                  'g_lamb', 'g_weight')] + \
               [('B', C.c_int), ('K', C.c_int), ('R', C.c_int), ('C', C.c_int), ('warp', C.c_int),
                # fused reconstruction-loss fields (device kernels only; unused by the host check)
                ('env_gt', FP), ('seg', FP), ('stats', C.POINTER(C.c_double)), ('g_num', FP), ('offset', C.c_float)]


class RenderEnvArgs(C.Structure):
    _fields_ = [(n, FP) for n in ('albedo', 'normal', 'rough', 'env', 'view', 'out_diff', 'out_spec', 'g_diff',
                                  'g_spec', 'g_albedo', 'g_normal', 'g_rough', 'g_env')] + \
               [('B', C.c_int), ('R', C.c_int), ('C', C.c_int), ('F0', C.c_float)]


def _needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    srcs = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh')]
    return any(os.path.getmtime(s) > t for s in srcs)


def load():
    if _needs_build():
        subprocess.check_call(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-o', LIB, SRC])
    lib = C.CDLL(LIB)
    assert lib.hc_sizeof_envtab() == C.sizeof(EnvTab)
    lib.hc_sizeof_render_sg_args.restype = C.c_size_t
    assert lib.hc_sizeof_render_sg_args() == C.sizeof(RenderSgArgs)
    for fn, st in ((lib.hc_sizeof_sg_env_args, SgEnvArgs), (lib.hc_sizeof_render_env_args, RenderEnvArgs)):
        fn.restype = C.c_size_t
        assert fn() == C.sizeof(st), st.__name__
    return lib


def make_tab(ls, ew):
    t = EnvTab()
    D = ls.shape[0]
    for d in range(D):
        t.lx[d], t.ly[d], t.lz[d], t.w[d] = float(ls[d, 0]), float(ls[d, 1]), float(ls[d, 2]), float(ew[d])
    t.D = D
    return t


def p(a):
    if a is None:
        return None
    assert a.dtype == np.float32 and a.flags['C_CONTIGUOUS']
    return a.ctypes.data_as(FP)
