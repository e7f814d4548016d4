"""NumPy-in / NumPy-out facade with the reference's ``BilateralGrid.py`` entry points, backed by the
CUDA solver (``irb_bg_*``): ``BilateralGrid(im, sigma_spatial, sigma_luma, sigma_chroma)`` with its
``splat / slice / blur / filter`` operators (BilateralGrid.py:43-107), ``bistochastize`` (:109-120),
``BilateralSolver(grid, params).solve / .solveGrad`` (:122-184) and the module-level
``solve(grid, target, confidence, bs_params, im_shape)`` / ``solveForGrad(grid, grad_f_x, confidence,
bs_params, im_shape, yhat, target)`` (:186-207) that ``BilateralLayer.BilateralFunction`` calls per image
in the reference.  Arrays are copied to the current CUDA device and back; there is no CPU path.
``im`` is rounded to float32 (the reference's caller passes float32, BilateralLayer.py:52).
``bistochastize`` returns the vectors ``(n, m)`` wrapped in ``scipy.sparse.diags`` like the reference."""
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


    # -- operators of BilateralGrid.py:90-107 (fp64 NumPy in / out) ---------------------------------------
    def _apply(self, op, x, rows_in):
        x = np.asarray(x, dtype=np.float64)
        flat = x.ndim == 1
        x2 = x.reshape(x.shape[0], -1)
        if x2.shape[0] != rows_in:
            raise ValueError('expected %d rows (got %s)' % (rows_in, x.shape))
        ch = x2.shape[1]
        outs = []
        for c0 in range(0, ch, 4):           # the workspace layout holds up to 4 channels
            xc = x2[:, c0:c0 + 4]
            k = xc.shape[1]
            ws, V = self._build(k)
            buf = torch.zeros(1, k, self.npixels, dtype=torch.float64, device=self._dev)
            buf[0, :, :rows_in] = torch.from_numpy(np.ascontiguousarray(xc.T)).to(self._dev)
            out = torch.empty_like(buf)
            with torch.cuda.device(self._dev):
                _lib.check(_lib.lib().irb_bg_apply(_lib.ptr(ws), op, _lib.ptr(buf), _lib.ptr(out), 1, k, self.H, self.W,
                                                   _lib.stream_ptr(self._dev)), 'irb_bg_apply')
            rows_out = self.npixels if op == 1 else V
            outs.append(out[0, :, :rows_out].t().contiguous().cpu().numpy())
        res = np.concatenate(outs, axis=1)
        return res[:, 0] if flat else res.reshape((res.shape[0],) + x.shape[1:])

    def splat(self, x):
        """S x: per-pixel values [P(,ch)] -> per-vertex sums [V(,ch)] (BilateralGrid.py:90-91)"""
        return self._apply(0, x, self.npixels)

    def slice(self, y):
        """S^T y: per-vertex values -> per-pixel values (BilateralGrid.py:93-94)"""
        return self._apply(1, y, self.nvertices)

    def blur(self, x):
        """2 dim x + sum of the five [1 0 1] neighbour sums (BilateralGrid.py:96-102)"""
        assert np.asarray(x).shape[0] == self.nvertices
        return self._apply(2, x, self.nvertices)

    def filter(self, x):
        """bilateral filter slice(blur(splat x)) / slice(blur(splat 1)) (BilateralGrid.py:104-107)"""
        x = np.asarray(x, dtype=np.float64)
        return self.slice(self.blur(self.splat(x))) / self.slice(self.blur(self.splat(np.ones_like(x))))

    def bistochastization(self):
        """(n[V], m[V]) of BilateralGrid.bistochastize, computed on the device when the grid was built"""
        ws, V = self._build(1)
        n = torch.empty(self.npixels, dtype=torch.float64, device=self._dev)
        m = torch.empty_like(n)
        with torch.cuda.device(self._dev):
            _lib.check(_lib.lib().irb_bg_bistochastization(_lib.ptr(ws), _lib.ptr(n), _lib.ptr(m), 1, self.H, self.W,
                                                           _lib.stream_ptr(self._dev)), 'irb_bg_bistochastization')
        return n[:V].cpu().numpy(), m[:V].cpu().numpy()

    @property
    def range_ok(self):
        """False if a grid coordinate left [0, 255): the guide was not in [0, 255] (the reference's hash requirement)"""
        ws, _ = self._build(1)
        fl = torch.empty(1, dtype=torch.int32, device=self._dev)
        with torch.cuda.device(self._dev):
            _lib.check(_lib.lib().irb_bg_range_flags(_lib.ptr(ws), _lib.ptr(fl), 1, self.H, self.W, _lib.stream_ptr(self._dev)),
                       'irb_bg_range_flags')
        return int(fl.item()) == 0

    @property
    def nlinks(self):
        ws, _ = self._build(1)
        nv = torch.empty(1, dtype=torch.int32, device=self._dev)
        nl = torch.empty(1, dtype=torch.int64, device=self._dev)
        with torch.cuda.device(self._dev):
            _lib.check(_lib.lib().irb_bg_stats(_lib.ptr(ws), _lib.ptr(nv), _lib.ptr(nl), 1, self.H, self.W,
                                               _lib.stream_ptr(self._dev)), 'irb_bg_stats')
        return int(nl.item())


def bistochastize(grid, maxiter=10):
    """BilateralGrid.py:109-120: diagonal matrices (Dn, Dm).  The device build runs the reference's 10 iterations."""
    if maxiter != 10:
        raise ValueError('the device build runs the reference default of 10 bistochastisation iterations')
    from scipy.sparse import diags
    n, m = grid.bistochastization()
    return diags(n, 0), diags(m, 0)


class BilateralSolver(object):
    """BilateralGrid.py:122-184: ``solve(x[P,ch], w[P,1]) -> (xhat[P,ch], yhat[V,ch])`` and
    ``solveGrad(x, w, saved_yhat, saved_target) -> (grad_t[P,ch], grad_conf[P,ch])`` in fp64 NumPy."""

    def __init__(self, grid, params):
        self.grid = grid
        self.params = params
        self.Dn, self.Dm = bistochastize(grid)

    def solve(self, x, w):
        x = np.asarray(x, dtype=np.float64)
        ch = x.shape[-1]
        xhat, yhat = solve(self.grid, x, np.asarray(w, dtype=np.float64), self.params, (self.grid.H, self.grid.W, ch))
        return xhat.reshape(-1, ch), yhat

    def solveGrad(self, x, w, saved_yhat, saved_target):
        x = np.asarray(x, dtype=np.float64)
        ch = x.shape[-1]
        gt, gc = _solve_for_grad(self.grid, x, w, self.params, (self.grid.H, self.grid.W, ch), saved_yhat, saved_target,
                                 per_channel_conf=True)
        return gt.reshape(-1, ch), gc


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


def _solve_for_grad(grid, grad_f_x, confidence, bs_params, im_shape, yhat, target, per_channel_conf=False):
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
    if not REQUIRES_CONF_GRAD:
        return grad_t, None
    if per_channel_conf:
        # BilateralSolver.solveGrad returns grad_conf per channel [P,ch] (BilateralGrid.py:179-181; the sum over channels is
        # taken by solveForGrad).  Channels are independent PCG solves, so each one is an exact single-channel call.
        if ch == 1:
            return grad_t, gc[0, 0].cpu().numpy().astype(np.float64).reshape(-1, 1)
        cols = [_solve_for_grad(grid, np.asarray(grad_f_x).reshape(-1, ch)[:, k:k + 1], confidence, bs_params,
                                (im_shape[0], im_shape[1], 1), np.asarray(yhat)[:, k:k + 1],
                                np.asarray(target).reshape(-1, ch)[:, k:k + 1], per_channel_conf=True)[1] for k in range(ch)]
        return grad_t, np.concatenate(cols, axis=1)
    grad_c = gc[0, 0].cpu().numpy().astype(np.float64).reshape(im_shape[0], im_shape[1])
    return grad_t, grad_c


def solveForGrad(grid, grad_f_x, confidence, bs_params, im_shape, yhat, target):
    """BilateralGrid.py:193-207.  Returns (grad_t reshaped to im_shape, grad_c [H,W])."""
    return _solve_for_grad(grid, grad_f_x, confidence, bs_params, im_shape, yhat, target)
