import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


def relerr(a, b):
    """inf-norm-normalised error ||a-b||_inf / ||b||_inf (SURVEY.md 8c parity metric)."""
    import numpy as np
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))


def parity_tol(ref32, ref64):
    """max(1e-4, 2*e_ref32), e_ref32 = the reference's (oracle's) own fp32 error against its fp64 run"""
    e = relerr(ref32, ref64)
    return max(1e-4, 2 * e) if e == e else 1e-4


def check_parity(ours, ref64, ref32, name=''):
    """assert ||ours - ref64||inf/||ref64||inf <= max(1e-4, 2*e_ref32); returns (e_ours, tol)"""
    e, tol = relerr(ours, ref64), parity_tol(ref32, ref64)
    assert e <= tol, (name, e, tol)
    return e, tol


RENDER_NAMES = ('diff', 'spec', 'g_albedo', 'g_normal', 'g_rough', 'g_axis', 'g_lamb', 'g_weight')


def oracle_render_both(ins64, gD, gS, loss=None, **kw):
    """the oracle's fused chain (output2env -> forwardEnv) + autograd in fp64 and in fp32 on the same inputs.
    Returns (ref64, ref32): lists [diff, spec, g_albedo, g_normal, g_rough, g_axis, g_lamb, g_weight] of NumPy arrays."""
    import torch
    from oracle import render_oracle as ro
    res = []
    for dt in (torch.float64, torch.float32):
        ins = [t.detach().to(dt).requires_grad_(True) for t in ins64]
        d, s = ro.render_sg(*ins, **kw)
        if loss is None:
            L = (d * gD.to(dt)).sum() + (s * gS.to(dt)).sum()
        else:
            L = loss(d, s)
        g = torch.autograd.grad(L, ins)
        res.append([d.detach().numpy(), s.detach().numpy()] + [x.numpy() for x in g])
    return res[0], res[1]
