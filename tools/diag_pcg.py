#!/usr/bin/env python
"""Per-iteration PCG scalars of the CUDA solver next to the oracle's recurrence (mode 4 debugging)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from inverserenderingofindoorscene_b200 import BilateralLayer as BL, _lib  # noqa: E402
from oracle import bilateral_oracle as bo  # noqa: E402

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 4
H, W = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (120, 160)
ch = 3 if mode == 0 else 1
sig, bsp = BL._PRESETS[mode]
guide = bo.synth_guide(1, H, W, blur_sigma=8.0, seed=1238 + mode)
rng = np.random.default_rng(mode)
pred = rng.random((1, ch, H, W)).astype(np.float32)
conf = (0.05 + 0.95 * rng.random((1, 1, H, W))).astype(np.float32)
# oracle with per-iteration trace
grid = bo.Grid(np.transpose(guide[0], (1, 2, 0)) * np.float32(255.0), sigma_luma=sig[0], sigma_chroma=sig[1], sigma_spatial=sig[2])
sol = bo.Solver(grid, bsp[0], bsp[1], bsp[2], int(bsp[3]))
x = np.transpose(pred[0], (1, 2, 0)).reshape(-1, ch).astype(np.double)
w = conf[0, 0].reshape(-1, 1).astype(np.double)
matvec, Minv, w_splat = sol._system(w)
b = grid.splat(x * w)
y0 = b / np.maximum(w_splat, 1e-10)[:, None]
print('V', grid.nvertices, 'P', H * W, 'links', grid.nlinks)
c = 0
xx = y0[:, c].copy()
r = b[:, c] - matvec(xx)
bb = float(b[:, c] @ b[:, c])
trace = []
p = None
rho_prev = None
for it in range(int(bsp[3])):
    rr = float(r @ r)
    z = Minv * r
    rho = float(r @ z)
    if np.sqrt(rr) < bsp[2] * np.sqrt(bb):
        trace.append((it, rho, None, rr, 'converged'))
        break
    p = z.copy() if it == 0 else p * (rho / rho_prev) + z
    q = matvec(p)
    pq = float(p @ q)
    trace.append((it, rho, pq, rr, ''))
    alpha = rho / pq
    xx = xx + alpha * p
    r = r - alpha * q
    rho_prev = rho
dev = torch.device('cuda')
lib = _lib.lib()
nb = lib.irb_bg_workspace_bytes(1, H, W, ch)
ws = torch.zeros(nb, dtype=torch.uint8, device=dev)
g = torch.from_numpy(guide).to(dev)
import ctypes as C
nv = (C.c_int * 1)()
st = _lib.stream_ptr(dev)
_lib.check(lib.irb_bg_build(_lib.ptr(g), 255.0, 1, H, W, float(sig[0]), float(sig[1]), float(sig[2]), _lib.ptr(ws), nb, nv, st), 'build')
out = torch.empty(1, ch, H, W, device=dev)
yhat = torch.zeros(1, ch, H * W, dtype=torch.float64, device=dev)
pred_d, conf_d = torch.from_numpy(pred).to(dev), torch.from_numpy(conf).to(dev)
_lib.check(lib.irb_bg_solve_fwd(_lib.ptr(ws), _lib.ptr(pred_d), _lib.ptr(conf_d), 1, ch, H, W,
                                float(bsp[0]), float(bsp[1]), float(bsp[2]), int(bsp[3]), _lib.ptr(out), _lib.ptr(yhat), st), 'solve')
sc = torch.empty(1, 4, 258, 4, dtype=torch.float64, device=dev)
_lib.check(lib.irb_bg_debug_scalars(_lib.ptr(ws), _lib.ptr(sc), 1, ch, H, W, st), 'dbg')
torch.cuda.synchronize()
sc = sc.cpu().numpy()
print('V gpu', nv[0], ' bb oracle %.6e gpu %.6e' % (bb, sc[0, c, 257, 0]))
for (it, rho, pq, rr, note) in trace:
    print('it %2d  rho %.6e / %.6e   pq %s / %.6e   rr %.6e / %.6e  %s' % (it, rho, sc[0, c, it, 0], ('%.6e' % pq) if pq else 'None', sc[0, c, it, 1], rr, sc[0, c, it, 2], note))
yh = yhat[0, c, :grid.nvertices].cpu().numpy()
print('yhat rel err', float(np.abs(yh - xx).max() / np.abs(xx).max()), ' y0 check', float(np.abs(y0[:, c]).max()))
