"""NumPy-in / NumPy-out facade with the reference's ``BilateralGrid.py`` entry points, backed by the
CUDA solver (``irb_bg_*``): ``BilateralGrid(im, sigma_spatial, sigma_luma, sigma_chroma)``,
``solve(grid, target, confidence, bs_params, im_shape)`` and ``solveForGrad(grid, grad_f_x,
confidence, bs_params, im_shape, yhat, target)`` (BilateralGrid.py:43-59, 186-207) -- the three calls
``BilateralLayer.BilateralFunction`` makes per image in the reference.  Arrays are copied to the
current CUDA device and back; there is no CPU path.  ``im`` is rounded to float32 (the reference's
caller passes float32, BilateralLayer.py:52)."""
import ctypes as C

import numpy as np
import torch

from . import _lib

REQUIRES_CONF_GRAD = True
MAX_VAL = 255.0


class BilateralGrid(object):
    def __init__(self, im, sigma_spatial=32, sigma_luma=8, sigma_chroma=8):
        im = np.ascontiguousarray(np.asarray(im, dtype=np.float32))
        if im.ndim != 3 or im.shape[2] != 3:
            raise ValueError('im must be [H,W,3]')
        self.H, self.W = im.shape[0], im.shape[1]
        self.npixels, self.dim = self.H * self.W, 5
        self.sigmas = (float(sigma_luma), float(sigma_chroma), float(sigma_spatial))
        self._dev = torch.device('cuda', torch.cuda.current_device())
        self._guide = torch.from_numpy(im).permute(2, 0, 1).contiguous().unsqueeze(0).to(self._dev)
        self._ws = {}
        self.nvertices = self._build(1)[1]

    def _build(self, ch):
        if ch not in self._ws:
            lib = _lib.lib()
            nbytes = lib.irb_bg_workspace_bytes(1, self.H, self.W, ch)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=self._dev)
            nv = (C.c_int * 1)()
            with torch.cuda.device(self._dev):
                _lib.check(lib.irb_bg_build(_lib.ptr(self._guide), 1.0, 1, self.H, self.W, self.sigmas[0], self.sigmas[1],
                                            self.sigmas[2], _lib.ptr(ws), nbytes, nv, _lib.stream_ptr(self._dev)),
                           'irb_bg_build')
            self._ws[ch] = (ws, int(nv[0]))
        return self._ws[ch]

    @property
    def idx(self):
        """vertex id of every pixel (rank of its hash among the sorted unique hashes)."""
        ws, _ = self._build(1)
        out = torch.empty(self.npixels, dtype=torch.int32, device=self._dev)
        with torch.cuda.device(self._dev):
            _lib.check(_lib.lib().irb_bg_vertex_index(_lib.ptr(ws), _lib.ptr(out), 1, self.H, self.W,
                                                      _lib.stream_ptr(self._dev)), 'irb_bg_vertex_index')
        return out.cpu().numpy().astype(np.int64)


def _params(bs_params):
    return (float(bs_params['lam']), float(bs_params['A_diag_min']), float(bs_params['cg_tol']),
            int(bs_params['cg_maxiter']))


def _to_dev(a, dev, ch):
    a = np.asarray(a, dtype=np.float32).reshape(-1, ch)
    return torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)  # [ch, P]


def solve(grid, target, confidence, bs_params, im_shape):
    """BilateralGrid.py:186-191.  Returns (xhat reshaped to im_shape (float64), yhat [V,ch] float64)."""
    ch = im_shape[2]
    lam, dmin, tol, mit = _params(bs_params)
    ws, V = grid._build(ch)
    dev = grid._dev
    t = _to_dev(target, dev, ch).unsqueeze(0)
    c = _to_dev(confidence, dev, 1).unsqueeze(0)
    out = torch.empty(1, ch, grid.npixels, dtype=torch.float32, device=dev)
    yhat = torch.empty(1, ch, grid.npixels, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().irb_bg_solve_fwd(_lib.ptr(ws), _lib.ptr(t), _lib.ptr(c), 1, ch, grid.H, grid.W, lam, dmin, tol,
                                               mit, _lib.ptr(out), _lib.ptr(yhat), _lib.stream_ptr(dev)), 'irb_bg_solve_fwd')
    xhat = out[0].t().contiguous().cpu().numpy().astype(np.float64).reshape(im_shape)
    return xhat, yhat[0, :, :V].t().contiguous().cpu().numpy()


def solveForGrad(grid, grad_f_x, confidence, bs_params, im_shape, yhat, target):
    """BilateralGrid.py:193-207.  Returns (grad_t reshaped to im_shape, grad_c [H,W])."""
    ch = im_shape[2]
    lam, dmin, tol, mit = _params(bs_params)
    ws, V = grid._build(ch)
    dev = grid._dev
    g = _to_dev(grad_f_x, dev, ch).unsqueeze(0)
    t = _to_dev(target, dev, ch).unsqueeze(0)
    c = _to_dev(confidence, dev, 1).unsqueeze(0)
    yh = torch.zeros(1, ch, grid.npixels, dtype=torch.float64, device=dev)
    yh[0, :, :V] = torch.from_numpy(np.ascontiguousarray(np.asarray(yhat, dtype=np.float64).T)).to(dev)
    gp = torch.empty(1, ch, grid.npixels, dtype=torch.float32, device=dev)
    gc = torch.empty(1, 1, grid.npixels, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().irb_bg_solve_bwd(_lib.ptr(ws), _lib.ptr(t), _lib.ptr(c), _lib.ptr(g), _lib.ptr(yh), 1, ch, grid.H,
                                               grid.W, lam, dmin, tol, mit, _lib.ptr(gp), _lib.ptr(gc), _lib.stream_ptr(dev)),
                   'irb_bg_solve_bwd')
    grad_t = gp[0].t().contiguous().cpu().numpy().astype(np.float64).reshape(im_shape)
    grad_c = gc[0, 0].cpu().numpy().astype(np.float64).reshape(im_shape[0], im_shape[1]) if REQUIRES_CONF_GRAD else None
    return grad_t, grad_c
