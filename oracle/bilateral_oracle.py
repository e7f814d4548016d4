"""CPU oracle (TEST INFRASTRUCTURE) for the bilateral-grid solver.

NumPy restatement -- not a copy -- of ``/root/reference/BilateralGrid.py`` and the
per-image bridge in ``/root/reference/BilateralLayer.py:20-124``.  The reference
builds explicit SciPy CSR matrices and calls ``scipy.sparse.linalg.cg``; here the
grid is a neighbour table ``nbr[V,10]`` and the PCG recurrence of SciPy 1.18
(``scipy/sparse/linalg/_isolve/iterative.py::cg``, the version in this image; the
reference pins none) is written out, with ``cg_maxiter`` an explicit integer
(the reference reads it from the wrong slot, ``BilateralLayer.py:44,79``; see
SURVEY.md 8c).  Pinned against the reference itself (with the two import shims of
SURVEY.md 8c) by ``oracle/make_golden.py`` -> ``tests/golden/bilateral_*.npz``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` may import this.
"""
import numpy as np

MAX_VAL = 255.0
# BilateralGrid.py:13-16 / 21
RGB_TO_YUV = np.array([[0.299, 0.587, 0.114],
                       [-0.168736, -0.331264, 0.5],
                       [0.5, -0.418688, -0.081312]])
YUV_OFFSET = np.array([0, 128.0, 128.0]).reshape(1, 1, -1)
DIM = 5


class Grid(object):
    """5-D XYLUV bilateral grid (BilateralGrid.py:44-88).

    ``im`` is [H,W,3] (the reference passes ``guide*255.0`` in float32,
    BilateralLayer.py:52).  Attributes: ``idx[P]`` vertex of each pixel (rank of its
    hash among the sorted unique hashes), ``nvertices``, ``coords[V,5]``,
    ``nbr[V,10]`` (column 2*d = coordinate d minus 1, 2*d+1 = plus 1; -1 if absent).
    """

    def __init__(self, im, sigma_spatial=32, sigma_luma=8, sigma_chroma=8):
        # BilateralGrid.py:23-24 -- np.tensordot -> fp64 dgemm
        im_yuv = np.tensordot(im, RGB_TO_YUV, ([2], [1])) + YUV_OFFSET
        H, W = im.shape[0], im.shape[1]
        Iy, Ix = np.mgrid[:H, :W]
        # BilateralGrid.py:47-52 -- truncation toward zero of fp64 quotients
        cols = [(Ix / sigma_spatial).astype(int), (Iy / sigma_spatial).astype(int),
                (im_yuv[..., 0] / sigma_luma).astype(int),
                (im_yuv[..., 1] / sigma_chroma).astype(int),
                (im_yuv[..., 2] / sigma_chroma).astype(int)]
        coords = np.stack(cols, axis=-1).reshape(-1, DIM)
        self.npixels = coords.shape[0]
        self.dim = DIM
        hash_vec = MAX_VAL ** np.arange(DIM)  # BilateralGrid.py:57
        hashed = coords.astype(np.float64) @ hash_vec  # BilateralGrid.py:86-88
        uniq, uidx, inv = np.unique(hashed, return_index=True, return_inverse=True)  # :64-65
        self.idx = inv.reshape(-1).astype(np.int64)
        self.coords = coords[uidx]
        self.nvertices = V = len(uniq)
        self.hashes = uniq
        # BilateralGrid.py:72-84 -- +-1 neighbours located by searchsorted
        nbr = np.full((V, 2 * DIM), -1, dtype=np.int64)
        for d in range(DIM):
            for s, off in enumerate((-1, 1)):
                nh = uniq + off * hash_vec[d]
                loc = np.clip(np.searchsorted(uniq, nh), 0, V - 1)
                ok = uniq[loc] == nh
                nbr[ok, 2 * d + s] = loc[ok]
        self.nbr = nbr
        self.nlinks = int((nbr >= 0).sum())

    # BilateralGrid.py:90-94
    def splat(self, x):
        x = np.asarray(x, dtype=np.float64)
        out = np.zeros((self.nvertices,) + x.shape[1:], dtype=np.float64)
        np.add.at(out, self.idx, x)
        return out

    def slice(self, y):
        return y[self.idx]

    # BilateralGrid.py:96-102 -- SUM of five [1 2 1] blurs
    def blur(self, x):
        out = 2 * DIM * x
        for j in range(2 * DIM):
            col = self.nbr[:, j]
            ok = col >= 0
            out[ok] = out[ok] + x[col[ok]]
        return out


def bistochastize(grid, maxiter=10):
    """BilateralGrid.py:109-120.  Returns (n[V], m[V])."""
    m = grid.splat(np.ones(grid.npixels))
    n = np.ones(grid.nvertices)
    for _ in range(maxiter):
        n = np.sqrt(n * m / grid.blur(n))
    m = n * grid.blur(n)
    return n, m


def _pcg(matvec, b, x0, Minv, maxiter, rtol):
    """SciPy 1.18 ``cg`` recurrence (see module docstring).  Returns (x, iterations)."""
    x = x0.copy()
    bnrm2 = np.linalg.norm(b)
    atol = max(0.0, float(rtol) * float(bnrm2))
    if bnrm2 == 0:
        return b.copy(), 0
    r = b - matvec(x) if x.any() else b.copy()
    rho_prev, p = None, None
    for it in range(maxiter):
        if np.linalg.norm(r) < atol:
            return x, it
        z = Minv * r
        rho = np.dot(r, z)
        if it > 0:
            p = p * (rho / rho_prev) + z
        else:
            p = z.copy()
        q = matvec(p)
        alpha = rho / np.dot(p, q)
        x = x + alpha * p
        r = r - alpha * q
        rho_prev = rho
    return x, maxiter


class Solver(object):
    """BilateralGrid.py:122-184 (BilateralSolver)."""

    def __init__(self, grid, lam, A_diag_min, cg_tol, cg_maxiter):
        self.grid = grid
        self.lam = float(lam)
        self.A_diag_min = float(A_diag_min)
        self.cg_tol = float(cg_tol)
        self.cg_maxiter = int(cg_maxiter)
        self.n, self.m = bistochastize(grid)

    def _system(self, w):
        g = self.grid
        n, m, lam = self.n, self.m, self.lam
        w_splat = g.splat(w)[:, 0]
        # A = lam*(Dm - Dn blur Dn) + diag(S w)   (BilateralGrid.py:134-137)
        diag = lam * (m - n * (2 * DIM) * n) + w_splat
        def matvec(x):
            return lam * (m * x - n * g.blur(n * x)) + w_splat * x
        Minv = 1.0 / np.maximum(diag, self.A_diag_min)  # :141-142
        return matvec, Minv, w_splat

    def solve(self, x, w):
        """x[P,ch], w[P,1] fp64 -> (xhat[P,ch], yhat[V,ch], iters[ch]).  BilateralGrid.py:128-150."""
        g = self.grid
        matvec, Minv, w_splat = self._system(w)
        b = g.splat(x * w)
        y0 = b / np.maximum(w_splat, 1e-10)[:, None]
        yhat = np.empty_like(y0)
        iters = []
        for d in range(x.shape[-1]):
            yhat[:, d], it = _pcg(matvec, b[:, d], y0[:, d], Minv, self.cg_maxiter, self.cg_tol)
            iters.append(it)
        return g.slice(yhat), yhat, iters

    def solve_grad(self, gx, w, saved_yhat, saved_target):
        """Adjoint solve, BilateralGrid.py:152-184.  Returns (grad_t[P,ch], grad_conf[P,ch])."""
        g = self.grid
        matvec, Minv, _ = self._system(w)
        b = g.splat(gx)
        ones = g.splat(np.ones((g.npixels, 1)))
        y0 = b / ones
        gh = np.empty_like(y0)
        for d in range(gx.shape[-1]):
            gh[:, d], _ = _pcg(matvec, b[:, d], y0[:, d], Minv, self.cg_maxiter, self.cg_tol)
        sl = g.slice(gh)
        grad_t = sl * w
        grad_conf = g.slice(-1.0 * (gh * saved_yhat)) + sl * saved_target
        return grad_t, grad_conf


def bilateral_forward(image, pred, conf, grid_params, bs_params):
    """BilateralFunction.forward (BilateralLayer.py:24-63) on NumPy arrays.

    image [B,3,H,W], pred [B,ch,H,W], conf [B,1,H,W] float32;
    grid_params=(sigma_luma, sigma_chroma, sigma_spatial); bs_params=(lam, A_diag_min,
    cg_tol, cg_maxiter:int).  Returns (out [B,ch,H,W] float32, yhat list, grids).
    """
    B, ch, H, W = pred.shape
    out = np.zeros((B, H, W, ch), np.float32)
    img = np.transpose(image, (0, 2, 3, 1))
    prd = np.transpose(pred, (0, 2, 3, 1))
    yhats, grids = [], []
    for i in range(B):
        grid = Grid(img[i] * 255.0, sigma_luma=grid_params[0], sigma_chroma=grid_params[1],
                    sigma_spatial=grid_params[2])
        sol = Solver(grid, *bs_params)
        xhat, yhat, _ = sol.solve(prd[i].reshape(-1, ch).astype(np.double),
                                  conf[i, 0].reshape(-1, 1).astype(np.double))
        out[i] = xhat.reshape(H, W, ch)
        yhats.append(yhat)
        grids.append(grid)
    return np.ascontiguousarray(np.transpose(out, (0, 3, 1, 2))), yhats, grids


def bilateral_backward(image, pred, conf, grad_out, yhats, grid_params, bs_params):
    """BilateralFunction.backward (BilateralLayer.py:66-124).  Returns (g_pred, g_conf) float32."""
    B, ch, H, W = pred.shape
    gp = np.zeros((B, H, W, ch), np.float32)
    gc = np.zeros((B, H, W), np.float32)
    img = np.transpose(image, (0, 2, 3, 1))
    prd = np.transpose(pred, (0, 2, 3, 1))
    go = np.transpose(grad_out, (0, 2, 3, 1))
    for i in range(B):
        grid = Grid(img[i] * 255.0, sigma_luma=grid_params[0], sigma_chroma=grid_params[1],
                    sigma_spatial=grid_params[2])
        sol = Solver(grid, *bs_params)
        gt, gcf = sol.solve_grad(go[i].reshape(-1, ch).astype(np.double),
                                 conf[i, 0].reshape(-1, 1).astype(np.double),
                                 yhats[i], prd[i].reshape(-1, ch).astype(np.double))
        gp[i] = gt.reshape(H, W, ch)
        gc[i] = gcf.reshape(H, W, ch).sum(2)
    return (np.ascontiguousarray(np.transpose(gp, (0, 3, 1, 2))), gc[:, None])


def synth_guide(B, H, W, blur_sigma=8.0, seed=1238):
    """Gaussian-blurred uniform noise, min-max normalised per image to [0,1].
    Separable blur written with NumPy only (no scipy.ndimage dependency)."""
    rng = np.random.default_rng(seed)
    x = rng.random((B, 3, H, W))
    if blur_sigma > 0:
        rad = int(4 * blur_sigma + 0.5)
        t = np.arange(-rad, rad + 1)
        kern = np.exp(-0.5 * (t / blur_sigma) ** 2)
        kern /= kern.sum()
        xp = np.pad(x, ((0, 0), (0, 0), (rad, rad), (0, 0)), mode='reflect')
        x = sum(kern[i] * xp[:, :, i:i + H, :] for i in range(2 * rad + 1))
        xp = np.pad(x, ((0, 0), (0, 0), (0, 0), (rad, rad)), mode='reflect')
        x = sum(kern[i] * xp[:, :, :, i:i + W] for i in range(2 * rad + 1))
    mn = x.reshape(B, -1).min(1).reshape(B, 1, 1, 1)
    mx = x.reshape(B, -1).max(1).reshape(B, 1, 1, 1)
    return ((x - mn) / (mx - mn)).astype(np.float32)
