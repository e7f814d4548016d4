"""Generate tests/golden/*.npz from the UNMODIFIED reference (build container only).

Run:  python oracle/make_golden.py            (needs /root/reference; CPU only)

The reference ships no golden vectors, so the fixtures are outputs of
the reference's own code on seeded synthetic inputs:
  * render_*.npz     -- models.output2env.output2env + models.renderingLayer.forwardEnv
                        (+ autograd) run in fp32 and, with the layer constants cast to
                        double, in fp64; LSregress / LSregressDiffSpec.
  * bilateral_*.npz  -- BilateralGrid.BilateralGrid / solve / solveForGrad with the two
                        import shims of SURVEY.md 8c (empty ``pylab`` module;
                        ``cg(tol=)`` -> ``cg(rtol=)``) and an explicit integer cg_maxiter.
Nothing from /root/reference is copied; only numeric inputs/outputs are stored.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = os.environ.get('IRB_REFERENCE', '/root/reference')
GOLD = os.path.join(ROOT, 'tests', 'golden')

from oracle import render_oracle as ro  # noqa: E402
from oracle import bilateral_oracle as bo  # noqa: E402


def load_reference():
    sys.path.insert(0, REF)
    import models as ref_models  # noqa
    sys.modules.setdefault('pylab', types.ModuleType('pylab'))
    import scipy.sparse.linalg as sla
    import BilateralGrid as ref_bg  # noqa
    _cg = sla.cg

    def cg_shim(A, b, x0=None, M=None, maxiter=None, tol=1e-5):
        return _cg(A, b, x0=x0, M=M, maxiter=int(maxiter), rtol=tol)
    ref_bg.cg = cg_shim
    return ref_models, ref_bg


def run_render(ref, dtype, ins, R, C, K, gD, gS, fov=57.0):
    alb, nrm, rg, ax, lam, wt = [t.clone().to(dtype).requires_grad_(True) for t in ins]
    rl = ref.renderingLayer(imWidth=C, imHeight=R, fov=fov, isCuda=False)
    o2e = ref.output2env(K, isCuda=False)
    if dtype == torch.float64:
        for a in ('v', 'ls', 'envWeight', 'up', 'pCoord'):
            setattr(rl, a, getattr(rl, a).double())
        o2e.ls = o2e.ls.double()
    env, _, lam_w, wt_w = o2e.output2env(ax, lam, wt)
    d, s = rl.forwardEnv(alb, nrm, rg, env)
    loss = (d * gD.to(dtype)).sum() + (s * gS.to(dtype)).sum()
    g = torch.autograd.grad(loss, [alb, nrm, rg, ax, lam, wt, env], retain_graph=True)
    out = dict(env=env, lam_w=lam_w, wt_w=wt_w, diff=d, spec=s, g_albedo=g[0], g_normal=g[1], g_rough=g[2],
               g_axis=g[3], g_lamb=g[4], g_weight=g[5], g_env=g[6])
    return {k: v.detach().numpy() for k, v in out.items()}, rl


def make_render(ref, name, B, H, W, R, C, K, seed, hi=1.0, fov=57.0, unit_normals=False):
    ins = ro.synth_inputs(B, H, W, R, C, K, seed=seed, dtype=torch.float32, hi=hi)
    if unit_normals:
        ins[1] = (ins[1] / ins[1].norm(dim=1, keepdim=True)).contiguous()
    g = torch.Generator().manual_seed(seed + 77)
    gD = torch.rand(B, 3, R, C, generator=g)
    gS = torch.rand(B, 3, R, C, generator=g)
    im = torch.rand(B, 3, R, C, generator=g)
    seg = (torch.rand(B, 1, R, C, generator=g) > 0.2).float()
    o64, rl = run_render(ref, torch.float64, ins, R, C, K, gD, gS, fov)
    o32, _ = run_render(ref, torch.float32, ins, R, C, K, gD, gS, fov)
    # regressions on the fp64 outputs (models.py:7-84)
    d64 = torch.from_numpy(o64['diff'])
    s64 = torch.from_numpy(o64['spec'])
    dS, sS = ref.LSregressDiffSpec(d64, s64, im.double(), d64, s64)
    e64 = torch.from_numpy(o64['env'])
    eS = ref.LSregress(e64 * 0.5, e64 + 0.1, e64)
    save = dict(albedo=ins[0].numpy(), normal=ins[1].numpy(), rough=ins[2].numpy(), axis=ins[3].numpy(),
                lamb=ins[4].numpy(), weight=ins[5].numpy(), gD=gD.numpy(), gS=gS.numpy(), im=im.numpy(),
                seg=seg.numpy(), meta=np.array([B, H, W, R, C, K], dtype=np.int64), fov=np.float64(fov),
                v=rl.v.numpy().astype(np.float32), lsreg_diff=dS.numpy(), lsreg_spec=sS.numpy(),
                lsreg_env=eS.numpy())
    for k, v in o64.items():
        save['ref64_' + k] = v
    for k, v in o32.items():
        save['ref32_' + k] = v
    path = os.path.join(GOLD, 'render_%s.npz' % name)
    np.savez_compressed(path, **save)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


def make_bilateral(ref_bg, name, H, W, ch, sig, bsp, seed, blur_sigma, B=1):
    guide = bo.synth_guide(B, H, W, blur_sigma=blur_sigma, seed=seed)
    rng = np.random.default_rng(seed + 5)
    pred = rng.random((B, ch, H, W)).astype(np.float32)
    conf = (0.05 + 0.95 * rng.random((B, 1, H, W))).astype(np.float32)
    gout = rng.standard_normal((B, ch, H, W)).astype(np.float32)
    sl, sc, ss = sig
    lam, dmin, tol, mit = bsp
    params = {'lam': lam, 'A_diag_min': dmin, 'cg_tol': tol, 'cg_maxiter': int(mit)}
    save = dict(guide=guide, pred=pred, conf=conf, gout=gout, sig=np.array(sig, np.float64),
                bsp=np.array(bsp, np.float64))
    for i in range(B):
        im = np.transpose(guide[i], (1, 2, 0)) * 255.0  # BilateralLayer.py:52 (float32 product)
        grid = ref_bg.BilateralGrid(im, sigma_luma=sl, sigma_chroma=sc, sigma_spatial=ss)
        S = grid.S.tocsc()
        idx = S.indices.astype(np.int64)  # one nonzero per column (pixel) -> its vertex
        cur_pred = np.transpose(pred[i], (1, 2, 0))
        cur_conf = conf[i, 0]
        xhat, yhat = ref_bg.solve(grid, cur_pred, cur_conf, params, cur_pred.shape)
        grid2 = ref_bg.BilateralGrid(im, sigma_luma=sl, sigma_chroma=sc, sigma_spatial=ss)
        gt, gc = ref_bg.solveForGrad(grid2, np.transpose(gout[i], (1, 2, 0)), cur_conf, params,
                                     cur_pred.shape, yhat, cur_pred)
        Dn, Dm = ref_bg.bistochastize(grid)
        ones = np.ones(grid.nvertices)
        save.update({'idx%d' % i: idx, 'nverts%d' % i: np.int64(grid.nvertices),
                     'xhat%d' % i: xhat, 'yhat%d' % i: yhat, 'g_pred%d' % i: gt, 'g_conf%d' % i: gc,
                     'n%d' % i: Dn.diagonal(), 'm%d' % i: Dm.diagonal(), 'blur1_%d' % i: grid.blur(ones)})
    path = os.path.join(GOLD, 'bilateral_%s.npz' % name)
    np.savez_compressed(path, **save)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


def make_shading(name, R, C, K, seed, hi=1.0):
    """utils.predToShading (utils.py:156-195) on a seeded SG volume.  ``utils`` imports h5py (absent here, never used
    on this path): an empty stub module stands in for it."""
    sys.path.insert(0, REF)
    sys.modules.setdefault('h5py', types.ModuleType('h5py'))
    import utils as ref_utils  # noqa
    _, _, _, axis, lamb, weight = ro.synth_inputs(1, R, C, R, C, K, seed=seed, dtype=torch.float32, hi=hi)
    pred = torch.cat([axis.reshape(1, 3 * K, R, C), lamb, weight], dim=1).numpy().astype(np.float32)
    shading = ref_utils.predToShading(pred, SGNum=K)
    np.savez_compressed(os.path.join(GOLD, 'shading_%s.npz' % name), pred=pred, shading=shading,
                        meta=np.array([R, C, K, 32, 16]))
    print('shading_%s' % name, shading.shape, shading.dtype, float(shading.max()))


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    if '--only-shading' in sys.argv:   # later addition: does not touch the committed render / bilateral fixtures
        make_shading('k12', R=6, C=8, K=12, seed=1243, hi=0.9)
        make_shading('k24_full', R=5, C=7, K=24, seed=1244, hi=1.0)
        return
    ref, ref_bg = load_reference()
    # render: 2x pooling (trainLight geometry), full-domain and typical SG ranges
    make_render(ref, 'pool2_full', B=2, H=12, W=16, R=6, C=8, K=12, seed=1235, hi=1.0)
    make_render(ref, 'pool2_typ', B=2, H=12, W=16, R=6, C=8, K=12, seed=1236, hi=0.9)
    # identity pooling with exactly-unit normals: exercises the upper clamp regime
    make_render(ref, 'ident_unit', B=1, H=6, W=8, R=6, C=8, K=12, seed=1237, hi=0.9, unit_normals=True)
    # non-integer adaptive pooling + other fov (testReal.py:320-346) and K=24 (config 5)
    make_render(ref, 'ragged_k24', B=1, H=13, W=19, R=5, C=7, K=24, seed=1238, hi=0.95, fov=42.75)
    # bilateral: modes 0 / 2 / 4 presets (BilateralLayer.py:131-189)
    make_bilateral(ref_bg, 'mode0', 48, 64, 3, (8, 2, 7), (200, 1e-5, 1e-5, 12), seed=1239, blur_sigma=3.0, B=2)
    make_bilateral(ref_bg, 'mode2', 40, 56, 1, (8, 2, 8), (300, 1e-5, 1e-5, 10), seed=1240, blur_sigma=3.0)
    make_bilateral(ref_bg, 'mode4', 40, 56, 1, (4, 2, 4), (100, 1e-5, 1e-5, 10), seed=1241, blur_sigma=2.0)
    make_bilateral(ref_bg, 'noise', 24, 32, 3, (8, 2, 7), (200, 1e-5, 1e-5, 12), seed=1242, blur_sigma=0.0)
    make_shading('k12', R=6, C=8, K=12, seed=1243, hi=0.9)
    make_shading('k24_full', R=5, C=7, K=24, seed=1244, hi=1.0)


if __name__ == '__main__':
    main()
