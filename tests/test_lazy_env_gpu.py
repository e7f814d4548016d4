"""GPU tests of the reference call pattern output2env.output2env -> renderingLayer.forwardEnv with the lazily
materialised env-map: no env-map is written when forwardEnv is its only consumer,
it is written once when the caller reads it, gradients of both consumers meet in the SG parameters, and any in-place
edit / gradient observer switches to the reference's exact semantics (unfused kernel on the env values)."""
import numpy as np
import pytest
import torch

from conftest import RENDER_NAMES, check_parity, oracle_render_both, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def irb():
    from inverserenderingofindoorscene_b200 import models, _lib
    _lib.lib()
    assert torch.cuda.is_available()
    return models


def _inputs(B=2, R=12, C=16, K=12, seed=321, hi=0.9):
    from oracle import render_oracle as ro
    return ro.synth_inputs(B, R, C, R, C, K, seed=seed, dtype=torch.float64, hi=hi)


def test_env_never_written_when_only_forwardEnv_reads_it(irb):
    """testReal.py:407-411 pattern: exactly two kernels (fused forward, fused backward), no sg_to_env launch"""
    from inverserenderingofindoorscene_b200 import _lib
    from inverserenderingofindoorscene_b200._lazy import LazyTensor
    B, R, C, K = 2, 12, 16, 12
    ins64 = _inputs(B, R, C, K)
    g = torch.Generator().manual_seed(1)
    gD, gS = torch.rand(B, 3, R, C, generator=g), torch.rand(B, 3, R, C, generator=g)
    ref64, ref32 = oracle_render_both(ins64, gD, gS)
    dev = torch.device('cuda')
    ins = [t.float().to(dev).requires_grad_(True) for t in ins64]
    rl = irb.renderingLayer(imWidth=C, imHeight=R, isCuda=True)
    o2e = irb.output2env(K, isCuda=True)
    lib = _lib.lib()
    n0 = lib.irb_launch_count()
    env, ax, lw, ww = o2e.output2env(ins[3], ins[4], ins[5])
    assert isinstance(env, LazyTensor) and tuple(env.shape) == (B, 3, R, C, 8, 16) and env.requires_grad
    assert ax is ins[3] and tuple(lw.shape) == (B, K, R, C) and tuple(ww.shape) == (B, 3 * K, R, C)
    d, s = rl.forwardEnv(ins[0], ins[1], ins[2], env)
    gg = torch.autograd.grad([d, s], ins, [gD.to(dev), gS.to(dev)])
    torch.cuda.synchronize()
    assert lib.irb_launch_count() - n0 == 2
    assert not env.irb_materialized and not lw.irb_materialized
    for t, r64, r32, name in zip([d, s] + list(gg), ref64, ref32, RENDER_NAMES):
        check_parity(t.detach().cpu().numpy(), r64, r32, name)
    # the warped lamb / weight results evaluate to the reference's literal expression on demand
    assert torch.allclose(lw + 0, torch.tan(np.pi / 2 * (0.999 * ins[4].detach())))
    assert lw.irb_materialized


def test_inplace_edit_of_env_renders_the_edited_values(irb):
    """env.mul_(0.5) between the two calls must render half the radiance (the reference would)"""
    B, R, C, K = 1, 8, 12, 12
    ins64 = _inputs(B, R, C, K, seed=5)
    dev = torch.device('cuda')
    ins = [t.float().to(dev) for t in ins64]
    rl = irb.renderingLayer(imWidth=C, imHeight=R, isCuda=True)
    o2e = irb.output2env(K, isCuda=True)
    env0, _, _, _ = o2e.output2env(ins[3], ins[4], ins[5])
    d0, s0 = rl.forwardEnv(ins[0], ins[1], ins[2], env0)
    env, _, _, _ = o2e.output2env(ins[3], ins[4], ins[5])
    env.mul_(0.5)
    assert not env.irb_untouched()
    d1, s1 = rl.forwardEnv(ins[0], ins[1], ins[2], env)
    assert relerr(d1.cpu().numpy(), (0.5 * d0).cpu().numpy()) <= 1e-5
    assert relerr(s1.cpu().numpy(), (0.5 * s0).cpu().numpy()) <= 1e-4
    # masked write through a view
    env, _, _, _ = o2e.output2env(ins[3], ins[4], ins[5])
    env.view(B, 3, R * C, 128)[:, :, : R * C // 2] = 0
    d2, s2 = rl.forwardEnv(ins[0], ins[1], ins[2], env)
    assert float(d2.view(B, 3, -1)[:, :, : R * C // 2].abs().max()) == 0
    assert relerr(d2.view(B, 3, -1)[:, :, R * C // 2:].cpu().numpy(), d0.view(B, 3, -1)[:, :, R * C // 2:].cpu().numpy()) <= 1e-5


def test_inplace_edit_of_sg_parameters(irb):
    """lobes modified in place after output2env(): a materialised env keeps the reference's semantics (rendered from the
    values it was computed from); an un-materialised one cannot and says so instead of rendering stale lighting"""
    B, R, C, K = 1, 8, 12, 12
    ins64 = _inputs(B, R, C, K, seed=6)
    dev = torch.device('cuda')
    ins = [t.float().to(dev) for t in ins64]
    rl = irb.renderingLayer(imWidth=C, imHeight=R, isCuda=True)
    o2e = irb.output2env(K, isCuda=True)
    d0, s0 = rl.forwardEnv(ins[0], ins[1], ins[2], o2e.output2env(ins[3], ins[4], ins[5])[0])
    wt = ins[5].clone()
    env, _, _, _ = o2e.output2env(ins[3], ins[4], wt)
    env.materialize()
    wt.mul_(0.0)
    d1, s1 = rl.forwardEnv(ins[0], ins[1], ins[2], env)      # unfused kernel on the env computed BEFORE the edit
    assert relerr(d1.cpu().numpy(), d0.cpu().numpy()) <= 1e-5
    wt = ins[5].clone()
    env, _, _, _ = o2e.output2env(ins[3], ins[4], wt)
    wt.mul_(0.0)
    with pytest.raises(RuntimeError, match='modified in place'):
        rl.forwardEnv(ins[0], ins[1], ins[2], env)


def test_hook_on_env_sees_the_render_gradient(irb):
    """register_hook / retain_grad on the env-map: the render gradient flows THROUGH env (unfused kernels) as in the reference"""
    B, R, C, K = 1, 8, 12, 12
    ins64 = _inputs(B, R, C, K, seed=7)
    dev = torch.device('cuda')
    ins = [t.float().to(dev).requires_grad_(True) for t in ins64]
    rl = irb.renderingLayer(imWidth=C, imHeight=R, isCuda=True)
    o2e = irb.output2env(K, isCuda=True)
    env, _, _, _ = o2e.output2env(ins[3], ins[4], ins[5])
    seen = []
    env.register_hook(lambda g: seen.append(g.detach().clone()))
    d, s = rl.forwardEnv(ins[0], ins[1], ins[2], env)
    (d.sum() + s.sum()).backward()
    assert len(seen) == 1 and tuple(seen[0].shape) == (B, 3, R, C, 8, 16) and float(seen[0].abs().max()) > 0
    g_hooked = [t.grad.clone() for t in ins[3:]]
    ins2 = [t.detach().clone().requires_grad_(True) for t in ins]
    d2, s2 = rl.forwardEnv(ins2[0], ins2[1], ins2[2], o2e.output2env(ins2[3], ins2[4], ins2[5])[0])
    (d2.sum() + s2.sum()).backward()
    for a, b in zip(g_hooked, [t.grad for t in ins2[3:]]):      # same gradients through either route
        assert relerr(a.cpu().numpy(), b.cpu().numpy()) <= 2e-4


def test_no_grad_and_fromSGtoIm(irb):
    B, R, C, K = 1, 8, 12, 12
    ins64 = _inputs(B, R, C, K, seed=8)
    dev = torch.device('cuda')
    ins = [t.float().to(dev).requires_grad_(True) for t in ins64]
    rl = irb.renderingLayer(imWidth=C, imHeight=R, isCuda=True)
    o2e = irb.output2env(K, isCuda=True)
    with torch.no_grad():
        env, _, _, _ = o2e.output2env(ins[3], ins[4], ins[5])
        assert not env.requires_grad
        d, s = rl.forwardEnv(ins[0], ins[1], ins[2], env)
        assert not d.requires_grad and not env.irb_materialized
    lam_w = torch.tan(np.pi / 2 * 0.999 * ins[4].detach())
    wt_w = torch.tan(np.pi / 2 * 0.999 * ins[5].detach())
    e2 = o2e.fromSGtoIm(ins[3].detach(), lam_w, wt_w)           # already-warped parameters (models.py:371-389)
    d2, s2 = rl.forwardEnv(ins[0].detach(), ins[1].detach(), ins[2].detach(), e2)
    assert not e2.irb_materialized
    assert relerr(d2.cpu().numpy(), d.cpu().numpy()) <= 1e-4 and relerr(s2.cpu().numpy(), s.cpu().numpy()) <= 1e-4
    assert relerr((e2 + 0).cpu().numpy(), (env + 0).cpu().numpy()) <= 1e-4
