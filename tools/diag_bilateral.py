#!/usr/bin/env python
"""Bilateral diagnostics on the GPU box: (1) do results depend on uninitialised workspace memory?  every preset solved with the
workspace pre-filled with 0x00 / 0xFF / 0x7F bytes and with the allocator's leftovers, compared with the NumPy oracle;
(2) PCG image-group size sweep (IRB_BG_GROUP) at config 4."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(poison):
    import numpy as np
    import torch
    from inverserenderingofindoorscene_b200 import BilateralLayer as BL
    from oracle import bilateral_oracle as bo
    dev = torch.device('cuda')
    junk = torch.rand(64 << 20, device=dev)   # dirty the allocator's free list
    del junk
    for mode, ch in ((0, 3), (2, 1), (4, 1)):
        sig, bsp = BL._PRESETS[mode]
        B, H, W = 2, 120, 160
        guide = bo.synth_guide(B, H, W, blur_sigma=8.0, seed=1238 + mode)
        rng = np.random.default_rng(mode)
        pred = rng.random((B, ch, H, W)).astype(np.float32)
        conf = (0.05 + 0.95 * rng.random((B, 1, H, W))).astype(np.float32)
        gout = rng.standard_normal((B, ch, H, W)).astype(np.float32)
        bsp_i = (bsp[0], bsp[1], bsp[2], int(bsp[3]))
        o_out, yh, _ = bo.bilateral_forward(guide, pred, conf, sig, bsp_i)
        o_gp, o_gc = bo.bilateral_backward(guide, pred, conf, gout, yh, sig, bsp_i)
        for rep in range(3):
            p = torch.from_numpy(pred).to(dev).requires_grad_(True)
            c = torch.from_numpy(conf).to(dev).requires_grad_(True)
            out, _ = BL.BilateralFunction.apply(torch.from_numpy(guide).to(dev), p, c,
                                                torch.tensor([float(x) for x in sig], device=dev),
                                                torch.tensor([float(x) for x in bsp], device=dev))
            gp, gc = torch.autograd.grad(out, [p, c], torch.from_numpy(gout).to(dev))
            rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
            print('poison=%s mode=%d rep=%d  out %.2e  g_pred %.2e  g_conf %.2e' % (
                poison, mode, rep, rel(out.detach().cpu().numpy(), o_out), rel(gp.cpu().numpy(), o_gp), rel(gc.cpu().numpy(), o_gc)),
                flush=True)


def sweep():
    import torch
    import bench_configs as bc
    from inverserenderingofindoorscene_b200 import BilateralLayer as BL
    r = bc.run_bilateral_mode0(BL, torch.device('cuda'), iters=5)
    print('group=%s fwd %.3f ms  fwd+bwd %.3f ms' % (os.environ.get('IRB_BG_GROUP', 'auto'), r['fwd_ms'], r['fwd_bwd_ms']), flush=True)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'child':
        child(os.environ.get('IRB_DEBUG_POISON', 'none'))
    elif len(sys.argv) > 1 and sys.argv[1] == 'sweep1':
        sweep()
    else:
        for poison in (None, '0x00', '0xFF', '0x7F'):
            env = dict(os.environ)
            env.pop('IRB_DEBUG_POISON', None)
            if poison:
                env['IRB_DEBUG_POISON'] = poison
            subprocess.run([sys.executable, __file__, 'child'], env=env)
        for g in ('1', '2', '3', '4', '8'):
            env = dict(os.environ)
            env['IRB_BG_GROUP'] = g
            subprocess.run([sys.executable, __file__, 'sweep1'], env=env)
