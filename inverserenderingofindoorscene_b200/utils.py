"""Host-side mirror of the hot-path helper in the reference's ``utils.py``.

``predToShading`` (utils.py:156-195, called by testReal.py:636-643 to write the shading PNGs): SG lighting
volume -> per-pixel diffuse irradiance over a 16 x 32 hemisphere table.  The reference evaluates it in NumPy on
the host (a [K,3,R,C,16,32] fp64 broadcast, ~0.9 GB of temporaries at 120 x 160); here it is one launch of
``irb_sg_shading`` (csrc/irb_shading.cu).

Same signature and return convention: a NumPy ``pred`` [1,7K,R,C] gives a NumPy float64 [3,R,C] (what the
reference returns; the values are computed in fp32 on the GPU).  A CUDA tensor [B,7K,R,C] is accepted as an
extension and gives a CUDA float32 tensor ([3,R,C] for B = 1 like the reference's squeeze, else [B,3,R,C]).
"""
import numpy as np
import torch

from . import _lib


def predToShading(pred, envWidth=32, envHeight=16, SGNum=12):
    as_numpy = isinstance(pred, np.ndarray)
    if as_numpy:
        if not torch.cuda.is_available():
            raise RuntimeError('predToShading runs on the GPU only: no CUDA device (this build has no CPU path)')
        t = torch.from_numpy(np.ascontiguousarray(pred, dtype=np.float32)).cuda()
    else:
        t = pred
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise ValueError('pred must be a NumPy array or a CUDA tensor: this build has no CPU path')
        if t.dtype != torch.float32:
            raise ValueError('pred must be float32 (got %s)' % t.dtype)
    if t.dim() != 4 or t.shape[1] != 7 * SGNum:
        raise ValueError('pred must be [B,%d,R,C] = cat(axis, lambOrig, weightOrig) (got %s)' % (7 * SGNum, tuple(t.shape)))
    if as_numpy and t.shape[0] != 1:
        raise ValueError('the reference squeezes a batch of one (utils.py:172); got B=%d' % t.shape[0])
    t = t.contiguous()
    B, _, R, C = t.shape
    out = torch.empty(B, 3, R, C, device=t.device, dtype=torch.float32)
    with torch.cuda.device(t.device):
        _lib.check(_lib.lib().irb_sg_shading(_lib.ptr(t), _lib.ptr(out), B, SGNum, R, C, int(envHeight), int(envWidth),
                                             _lib.stream_ptr(t.device)), 'irb_sg_shading')
    if as_numpy:
        return out[0].double().cpu().numpy()
    return out[0] if B == 1 else out
