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


# ------------------------------------------------------------------------------------------------------------
# Cascade hand-off formats: what outputBRDFLight.py writes and dataLoader.py reads around the render layer.
# Host I/O, not kernels: the layouts are mirrored here so that the drop-in layer's outputs go to disk exactly as the
# reference's do.  HDF5 itself is h5py's job (imported lazily; this image does not ship it, the reference's own
# environment does).
# ------------------------------------------------------------------------------------------------------------

def packSGVolume(axisPred, lambPred, weightPred):
    """the 7K-channel SG volume ``envmapsPred`` of wrapperBRDFLight.py:168-169 -- what ``imenv_*_{n}.h5`` stores
    (outputBRDFLight.py:290-301), what cascade 1 feeds to its light encoder (dataLoader.py:171-184) and what
    ``predToShading`` takes: cat(axis [B,K,3,R,C] viewed as [B,3K,R,C], lambOrig [B,K,R,C], weightOrig [B,3K,R,C])."""
    bn, K, three, R, C = axisPred.shape
    if three != 3 or tuple(lambPred.shape) != (bn, K, R, C) or tuple(weightPred.shape) != (bn, 3 * K, R, C):
        raise ValueError('expected axis [B,K,3,R,C], lamb [B,K,R,C], weight [B,3K,R,C]')
    return torch.cat([axisPred.reshape(bn, K * 3, R, C), lambPred, weightPred], dim=1)


def unpackSGVolume(envmapsPred, SGNum=12):
    """inverse of ``packSGVolume``: [B,7K,R,C] -> (axis [B,K,3,R,C], lambOrig [B,K,R,C], weightOrig [B,3K,R,C]) -- the
    arguments of ``output2env.output2env`` / ``renderingLayer.forwardSG``."""
    bn, ch, R, C = envmapsPred.shape
    K = SGNum
    if ch != 7 * K:
        raise ValueError('expected %d channels (got %d)' % (7 * K, ch))
    return (envmapsPred[:, :3 * K].reshape(bn, K, 3, R, C), envmapsPred[:, 3 * K:4 * K], envmapsPred[:, 4 * K:])


def _h5py():
    try:
        import h5py
    except ImportError as ex:   # pragma: no cover - depends on the environment
        raise RuntimeError('the cascade hand-off files are HDF5: install h5py (the reference imports it at utils.py:8, '
                           'dataLoader.py:10); this build image does not ship it') from ex
    return h5py


def writeH5ToFile(imBatch, nameBatch):
    """utils.py:92-99: one file per sample, dataset ``data`` = the sample's [ch,H,W] float array, LZF-compressed --
    the format of ``imdiffuse_* / imspecular_* / imenv_* / imbaseColor_* ...`` (outputBRDFLight.py:235-301)."""
    h5py = _h5py()
    if imBatch.size(0) != len(nameBatch):
        raise ValueError('one file name per sample')
    for n in range(imBatch.size(0)):
        im = imBatch[n].detach().cpu().numpy()
        with h5py.File(nameBatch[n], 'w') as hf:
            hf.create_dataset('data', data=im, compression='lzf')


def loadH5(imName):
    """dataLoader.py:277-283: the ``data`` dataset as a NumPy array, ``None`` if the file cannot be read"""
    h5py = _h5py()
    try:
        with h5py.File(imName, 'r') as hf:
            return np.array(hf.get('data'))
    except Exception:
        return None
