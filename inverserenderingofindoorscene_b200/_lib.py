"""ctypes binding of libirb_b200.so -- the C ABI declared in include/irb.h.

There is NO fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
# IRB_LIB selects another build of the same sources (A/B variants made by _build.build(variant=...)); default = the product
LIB_PATH = os.environ.get('IRB_LIB') or os.path.join(PKG, 'libirb_b200.so')

_P = C.c_void_p
_I = C.c_int
_F = C.c_float
_D = C.c_double

# name -> argtypes (restype int unless listed in _RESTYPE); mirrors include/irb.h one to one
SIGNATURES = {
    'irb_abi_version': [],
    'irb_last_error': [],
    'irb_launch_count': [],
    'irb_sg_to_env_fwd': [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    'irb_sg_to_env_bwd': [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    'irb_render_env_fwd': [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P],
    'irb_render_env_bwd': [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P],
    'irb_render_sg_fwd': [_P] * 10 + [_I, _I, _I, _I, _I, _I, _F, _I, _P],
    'irb_render_sg_bwd': [_P] * 16 + [_I, _I, _I, _I, _I, _I, _F, _I, _P],
    'irb_avgpool_fwd': [_P, _P, _I, _I, _I, _I, _I, _P],
    'irb_avgpool_bwd': [_P, _P, _I, _I, _I, _I, _I, _P],
    'irb_avgpool3_fwd': [_P] * 6 + [_I] * 7 + [_P],
    'irb_avgpool3_bwd': [_P] * 6 + [_I] * 7 + [_P],
    'irb_lsregress_diffspec': [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, C.c_longlong, _P],
    'irb_render_loss_fwd': [_P] * 9 + [_I, _I, C.c_longlong, _P],
    'irb_render_loss_bwd': [_P] * 5 + [_I, _I, C.c_longlong, _P],
    'irb_lsregress': [_P, _P, _P, _P, _P, _P, _I, C.c_longlong, _P],
    'irb_sg_reconst_fwd': [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _P],
    'irb_sg_reconst_bwd': [_P] * 10 + [_I, _I, _I, _I, _I, _I, _I, _F, _P],
    'irb_sg_shading': [_P, _P, _I, _I, _I, _I, _I, _I, _P],
    'irb_bg_workspace_bytes': [_I, _I, _I, _I],
    'irb_bg_build': [_P, _F, _I, _I, _I, _D, _D, _D, _P, C.c_size_t, C.POINTER(C.c_int), _P],
    'irb_bg_vertex_index': [_P, _P, _I, _I, _I, _P],
    'irb_bg_apply': [_P, _I, _P, _P, _I, _I, _I, _I, _P],
    'irb_bg_bistochastization': [_P, _P, _P, _I, _I, _I, _P],
    'irb_bg_stats': [_P, _P, _P, _I, _I, _I, _P],
    'irb_bg_range_flags': [_P, _P, _I, _I, _I, _P],
    'irb_bg_debug_scalars': [_P, _P, _I, _I, _I, _I, _P],
    'irb_bg_solve_fwd': [_P, _P, _P, _I, _I, _I, _I, _D, _D, _D, _I, _P, _P, _P],
    'irb_bg_solve_bwd': [_P, _P, _P, _P, _P, _I, _I, _I, _I, _D, _D, _D, _I, _P, _P, _P],
}
_RESTYPE = {'irb_last_error': C.c_char_p, 'irb_launch_count': C.c_longlong, 'irb_bg_workspace_bytes': C.c_size_t}

_lib = None


def lib():
    """Load (once) and return the CDLL.  Raises RuntimeError if the CUDA library is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                'libirb_b200.so is not built (%s). Run `python -m inverserenderingofindoorscene_b200._build` '
                '(or __graft_entry__.build()); there is no CPU fallback.' % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the ABI is incomplete
            fn.argtypes = args
            fn.restype = _RESTYPE.get(name, C.c_int)
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().irb_last_error()
        raise RuntimeError('%s failed (rc=%d): %s' % (what, rc, msg.decode() if msg else ''))


def ptr(t):
    """device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
