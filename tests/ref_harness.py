"""Harness that drives the reference's OWN caller, wrapperBRDFLight.wrapperBRDFLight (wrapperBRDFLight.py:8-223), with
the reference's own CNNs (randomly initialised: no checkpoints offline) on one synthetic dataBatch
(dataLoader.py:188-201 keys).  Test / bench infrastructure: imports the unmodified reference through
oracle/ref_loader.py (baseline/_ref)."""
import types

import torch


def make_opt(envRow=120, envCol=160, SGNum=12):
    return types.SimpleNamespace(cascadeLevel=0, envRow=envRow, envCol=envCol, envWidth=16, envHeight=8, SGNum=SGNum,
                                 imHeight=240, imWidth=320, cuda=True)


def make_batch(B, seed=0, H=240, W=320, envRow=120, envCol=160):
    g = torch.Generator().manual_seed(seed)
    n = torch.randn(B, 3, H, W, generator=g)
    n[:, 2] += 1.5
    n = n / n.norm(dim=1, keepdim=True)
    seg_obj = (torch.rand(B, 1, H, W, generator=g) > 0.15).float()
    seg_area = ((torch.rand(B, 1, H, W, generator=g) > 0.9).float() * (1 - seg_obj))
    seg_env = 1 - torch.clamp(seg_obj + seg_area, 0, 1)
    return {
        'albedo': torch.rand(B, 3, H, W, generator=g), 'normal': n, 'rough': torch.rand(B, 1, H, W, generator=g) * 2 - 1,
        'depth': torch.rand(B, 1, H, W, generator=g) * 3 + 0.5, 'segArea': seg_area, 'segEnv': seg_env, 'segObj': seg_obj,
        'im': torch.rand(B, 3, H, W, generator=g) * 0.95,
        'envmaps': torch.rand(B, 3, envRow, envCol, 8, 16, generator=g) * 2,
        'envmapsInd': torch.ones(B, 1, 1, 1),
    }


def make_nets(ref_models, SGNum=12, seed=0, device='cuda'):
    torch.manual_seed(seed)
    nets = dict(
        encoder=ref_models.encoder0(cascadeLevel=0), albedoDecoder=ref_models.decoder0(mode=0),
        normalDecoder=ref_models.decoder0(mode=1), roughDecoder=ref_models.decoder0(mode=2),
        depthDecoder=ref_models.decoder0(mode=4), lightEncoder=ref_models.encoderLight(cascadeLevel=0, SGNum=SGNum),
        axisDecoder=ref_models.decoderLight(mode=0, SGNum=SGNum), lambDecoder=ref_models.decoderLight(mode=1, SGNum=SGNum),
        weightDecoder=ref_models.decoderLight(mode=2, SGNum=SGNum))
    for k in ('encoder', 'albedoDecoder', 'normalDecoder', 'roughDecoder', 'depthDecoder'):   # frozen, trainLight.py:121-145
        for p in nets[k].parameters():
            p.requires_grad = False
    for m in nets.values():
        m.to(device)
    return nets


LIGHT = ('lightEncoder', 'axisDecoder', 'lambDecoder', 'weightDecoder')


def run_step(wrapper_mod, ref_models, nets, batch, opt, backward=True):
    """one trainLight.py:210-238 step: wrapper forward, totalErr = renderErr + 10 reconstErr, backward.
    Returns (reconstErr, renderErr, {param name: grad}) -- the render layer objects are built from whatever
    ref_models.renderingLayer / output2env currently ARE (reference or patched drop-in)."""
    renderLayer = ref_models.renderingLayer(isCuda=opt.cuda, imWidth=opt.envCol, imHeight=opt.envRow, envWidth=opt.envWidth,
                                            envHeight=opt.envHeight)
    output2env = ref_models.output2env(isCuda=opt.cuda, envWidth=opt.envWidth, envHeight=opt.envHeight, SGNum=opt.SGNum)
    for k in LIGHT:
        nets[k].zero_grad(set_to_none=True)
    res = wrapper_mod.wrapperBRDFLight(
        batch, opt, nets['encoder'], nets['albedoDecoder'], nets['normalDecoder'], nets['roughDecoder'], nets['depthDecoder'],
        nets['lightEncoder'], nets['axisDecoder'], nets['lambDecoder'], nets['weightDecoder'], output2env, renderLayer,
        offset=1.0)
    reconstErr, renderErr = res[4][1], res[5][1]
    grads = {}
    if backward:
        (1.0 * renderErr + 10.0 * reconstErr).backward()
        for k in LIGHT:
            for n, p in nets[k].named_parameters():
                grads[k + '.' + n] = None if p.grad is None else p.grad.detach().clone()
    return reconstErr.detach(), renderErr.detach(), grads, res
