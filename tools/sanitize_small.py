#!/usr/bin/env python
"""Small end-to-end invocations for compute-sanitizer (racecheck / initcheck / memcheck): the TMA-staged group kernels with the
persistent grid capped so that every CTA walks several tiles through the 2-stage ring, the unfused kernels, the fused losses and
one small bilateral solve.   compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault('IRB_TMA_MAX_CTAS', '2')
import torch  # noqa: E402

from inverserenderingofindoorscene_b200 import BilateralLayer as BL  # noqa: E402
from inverserenderingofindoorscene_b200 import models as M  # noqa: E402

dev = torch.device('cuda')
g = torch.Generator(device=dev).manual_seed(0)
B, R, C, K = 2, 16, 24, 12
alb = torch.rand(B, 3, R, C, device=dev, generator=g).requires_grad_(True)
n = torch.randn(B, 3, R, C, device=dev, generator=g)
n[:, 2] += 1.5
n = (n / n.norm(dim=1, keepdim=True) * 0.9).requires_grad_(True)
rg = (torch.rand(B, 1, R, C, device=dev, generator=g) * 2 - 1).requires_grad_(True)
ax = torch.randn(B, K, 3, R, C, device=dev, generator=g)
ax = (ax / ax.norm(dim=2, keepdim=True)).requires_grad_(True)
lam = torch.rand(B, K, R, C, device=dev, generator=g).requires_grad_(True)
wt = torch.rand(B, 3 * K, R, C, device=dev, generator=g).requires_grad_(True)
im = torch.rand(B, 3, R, C, device=dev, generator=g)
seg = (torch.rand(B, 1, R, C, device=dev, generator=g) > 0.2).float()
rl = M.renderingLayer(imWidth=C, imHeight=R, isCuda=True)
o2e = M.output2env(K, isCuda=True)
env, _, _, _ = o2e.output2env(ax, lam, wt)
d, s = rl.forwardEnv(alb, n, rg, env)                       # fused group kernels
_, err = M.renderLoss(d, s, im, seg)
envgt = torch.rand(B, 3, R, C, 8, 16, device=dev, generator=g)
rec, _ = o2e.reconstLoss(ax, lam, wt, envgt, seg)
d2, s2 = rl.forwardEnv(alb, n, rg, env * 1.0)               # sg_to_env + render_env
(err + rec + d2.sum() + s2.sum()).backward()
sig, bsp = BL._PRESETS[0]
guide = torch.rand(1, 3, 24, 32, device=dev, generator=g)
p = torch.rand(1, 3, 24, 32, device=dev, generator=g).requires_grad_(True)
c = (0.05 + 0.95 * torch.rand(1, 1, 24, 32, device=dev, generator=g)).requires_grad_(True)
out, _ = BL.BilateralFunction.apply(guide, p, c, torch.tensor([float(x) for x in sig], device=dev),
                                    torch.tensor([float(x) for x in bsp], device=dev))
out.sum().backward()
torch.cuda.synchronize()
print('sanitize_small done', float(err), float(rec), float(out.sum()))
