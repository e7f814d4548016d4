"""CPU oracle (TEST INFRASTRUCTURE) for the SG-lit microfacet render layer.

A restatement -- NOT a copy -- of the reference's PyTorch op graph
(``/root/reference/models.py``), written direction-by-direction on ``[B,R,C]``
planes instead of the reference's 5-7-D broadcast temporaries, so that it runs in
fp64 on a laptop-sized memory budget and autograd through it yields the "truth"
gradients the CUDA backward is checked against.  Every function cites the
reference lines it follows.  Pinned against outputs of the reference itself by
``oracle/make_golden.py`` -> ``tests/golden/render_*.npz``
(``tests/test_oracle_golden.py``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (cpu_baseline /
--impl reference) may import this module.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# constants (models.py:349-369 and models.py:408-459)
# --------------------------------------------------------------------------------------


def env_table(envHeight=8, envWidth=16):
    """Hemisphere direction table and solid-angle weights.

    models.py:353-362 / 437-448: Az_j=((j+.5)/W-.5)*2pi, El_i=((i+.5)/H)*pi/2,
    ls[d]=(sinEl cosAz, sinEl sinAz, cosEl) with d=i*W+j, envWeight[d]=sinEl*pi^2/W/H.
    Computed in fp64 and rounded to fp32 exactly as the reference does.
    Returns (ls[D,3] float32, envWeight[D] float32) numpy arrays.
    """
    Az = ((np.arange(envWidth) + 0.5) / envWidth - 0.5) * 2 * np.pi
    El = ((np.arange(envHeight) + 0.5) / envHeight) * np.pi / 2.0
    Az, El = np.meshgrid(Az, El)
    Az = Az.reshape(-1)
    El = El.reshape(-1)
    ls = np.stack([np.sin(El) * np.cos(Az), np.sin(El) * np.sin(Az), np.cos(El)], axis=1)
    w = np.sin(El) * np.pi * np.pi / envWidth / envHeight
    return ls.astype(np.float32), w.astype(np.float32)


def camera_rays(imHeight, imWidth, fov=57.0, cameraPos=(0.0, 0.0, 0.0)):
    """Per-pixel unit view vectors v[3,H,W] (float32), models.py:415-432.

    Pixel grid linspace(-xRange,xRange,W) x linspace(-yRange,yRange,H) with y flipped,
    z=-1; v = normalize(cameraPos - p) with the reference's fp32 arithmetic.
    """
    fov_r = fov / 180.0 * np.pi
    xRange = 1 * np.tan(fov_r / 2)
    yRange = float(imHeight) / float(imWidth) * xRange
    x, y = np.meshgrid(np.linspace(-xRange, xRange, imWidth), np.linspace(-yRange, yRange, imHeight))
    y = np.flip(y, axis=0)
    z = -np.ones((imHeight, imWidth), dtype=np.float32)
    p = np.stack([x, y, z]).astype(np.float32)
    cam = np.array(cameraPos, dtype=np.float32).reshape(3, 1, 1)
    v = cam - p
    v = v / np.sqrt(np.maximum(np.sum(v * v, axis=0), 1e-12))[None]
    return np.ascontiguousarray(v.astype(np.float32))


# --------------------------------------------------------------------------------------
# output2env (models.py:371-404)
# --------------------------------------------------------------------------------------

_WARP = 0.999


def sg_warp(lambOrig, weightOrig):
    """tan-warp of sharpness / intensity, models.py:396-400 (same op order)."""
    weight = torch.tan(np.pi / 2 * (_WARP * weightOrig))
    lamb = torch.tan(np.pi / 2 * (_WARP * lambOrig))
    return lamb, weight


def sg_to_env(axis, lamb, weight, envHeight=8, envWidth=16):
    """E[b,c,r,x,d] = sum_k W[b,k,c,r,x] exp(lam[b,k,r,x] (a[b,k,:,r,x].w_d - 1)).

    models.py:371-389 (fromSGtoIm).  axis [B,K,3,R,C], lamb [B,K,R,C] (already warped),
    weight [B,3K,R,C] with channel = k*3+c.  Returns [B,3,R,C,envH,envW].
    """
    B, K = axis.shape[0], axis.shape[1]
    R, C = axis.shape[3], axis.shape[4]
    ls, _ = env_table(envHeight, envWidth)
    ls = torch.from_numpy(ls).to(device=axis.device, dtype=axis.dtype)
    D = ls.shape[0]
    w = weight.reshape(B, K, 3, R, C)
    out = []
    for d in range(D):
        dot = axis[:, :, 0] * ls[d, 0] + axis[:, :, 1] * ls[d, 1] + axis[:, :, 2] * ls[d, 2]  # [B,K,R,C]
        e = torch.exp(lamb * (dot - 1))
        out.append((w * e.unsqueeze(2)).sum(dim=1))  # [B,3,R,C]
    env = torch.stack(out, dim=-1)
    return env.reshape(B, 3, R, C, envHeight, envWidth)


def output2env(axisOrig, lambOrig, weightOrig, envHeight=8, envWidth=16):
    """models.py:391-404: returns (envmaps, axis, lamb, weight)."""
    lamb, weight = sg_warp(lambOrig, weightOrig)
    return sg_to_env(axisOrig, lamb, weight, envHeight, envWidth), axisOrig, lamb, weight


# --------------------------------------------------------------------------------------
# renderingLayer.forwardEnv (models.py:461-522)
# --------------------------------------------------------------------------------------


def _normalize(x, dim, eps=1e-12):
    # F.normalize semantics: x / max(||x||, eps)
    return x / x.norm(dim=dim, keepdim=True).clamp_min(eps)


def render_env(albedo, normal, rough, envmap, fov=57.0, F0=0.05, cameraPos=(0.0, 0.0, 0.0), view=None):
    """colorDiffuse, colorSpec [B,3,R,C] of models.py:461-522.

    albedo/normal [B,3,H,W], rough [B,1,H,W] are pooled to the env grid R x C
    (models.py:465-469); envmap [B,3,R,C,envH,envW].  ``view`` optionally overrides the
    per-pixel view vectors [3,R,C] (used to render a crop of a larger frame).
    """
    dt = envmap.dtype
    B = albedo.shape[0]
    R, C = envmap.shape[2], envmap.shape[3]
    envH, envW = envmap.shape[4], envmap.shape[5]
    D = envH * envW
    ls_np, ew_np = env_table(envH, envW)
    dev = envmap.device
    ls = torch.from_numpy(ls_np).to(device=dev, dtype=dt)
    ew = torch.from_numpy(ew_np).to(device=dev, dtype=dt)
    v = (torch.from_numpy(camera_rays(R, C, fov, cameraPos)) if view is None else view).to(device=dev, dtype=dt)  # [3,R,C]

    # models.py:465-469
    A = F.adaptive_avg_pool2d(albedo, (R, C))
    n = F.adaptive_avg_pool2d(normal, (R, C))
    n = n / torch.sqrt(torch.clamp((n * n).sum(dim=1), 1e-6, 1).unsqueeze(1))
    rg = F.adaptive_avg_pool2d(rough, (R, C))

    # models.py:477-480: local frame
    up = torch.tensor([0.0, 1.0, 0.0], dtype=dt, device=dev).view(1, 3, 1, 1)
    upn = n[:, 1:2]  # up . n
    camy = _normalize(up - upn * n, dim=1)
    camx = -_normalize(torch.cross(camy, n, dim=1), dim=1)

    # models.py:493-500
    rho = (rg + 1.0) / 2.0
    k = (rho + 1) * (rho + 1) / 8.0
    alpha2 = (rho * rho) * (rho * rho)
    vv = v.unsqueeze(0)  # [1,3,R,C]
    ndv = torch.clamp((n * vv).sum(dim=1, keepdim=True), 0, 1)

    env = envmap.reshape(B, 3, R, C, D)
    cd = torch.zeros(B, 3, R, C, dtype=dt, device=dev)
    cs = torch.zeros(B, 3, R, C, dtype=dt, device=dev)
    two = torch.tensor(2.0, dtype=dt, device=dev)
    for d in range(D):
        # models.py:482-484
        l = ls[d, 0] * camx + ls[d, 1] * camy + ls[d, 2] * n
        # models.py:486-491
        h = (vv + l) / 2
        h = h / torch.sqrt(torch.clamp((h * h).sum(dim=1, keepdim=True), min=1e-6))
        vdh = (vv * h).sum(dim=1, keepdim=True)
        frac0 = F0 + (1 - F0) * torch.pow(two, (-5.55472 * vdh - 6.98316) * vdh)
        # models.py:501-509
        ndh = torch.clamp((n * h).sum(dim=1, keepdim=True), 0, 1)
        ndl = torch.clamp((n * l).sum(dim=1, keepdim=True), 0, 1)
        nom0 = ndh * ndh * (alpha2 - 1) + 1
        nom1 = ndv * (1 - k) + k
        nom2 = ndl * (1 - k) + k
        nom = torch.clamp(4 * np.pi * nom0 * nom0 * nom1 * nom2, 1e-6, 4 * np.pi)
        spec = alpha2 * frac0 / nom
        # models.py:511-520
        E = env[..., d]
        cd = cd + (A / np.pi) * ndl * E * ew[d]
        cs = cs + spec * ndl * E * ew[d]
    return cd, cs


def render_sg(albedo, normal, rough, axisOrig, lambOrig, weightOrig, fov=57.0, F0=0.05,
              cameraPos=(0.0, 0.0, 0.0), envHeight=8, envWidth=16):
    """output2env followed by forwardEnv (wrapperBRDFLight.py:177,194): the fused path."""
    env, _, _, _ = output2env(axisOrig, lambOrig, weightOrig, envHeight, envWidth)
    return render_env(albedo, normal, rough, env, fov, F0, cameraPos)


# --------------------------------------------------------------------------------------
# scale-ambiguity regressions and losses (models.py:7-84, wrapperBRDFLight.py:170-207)
# --------------------------------------------------------------------------------------


def ls_regress(pred, gt, origin):
    """models.py:7-21: per-sample scalar sum(p g)/clamp(sum p^2,1e-5), detached, clamp[1e-3,1e3]."""
    nb = pred.shape[0]
    p = pred.reshape(nb, -1)
    g = gt.reshape(nb, -1)
    coef = ((p * g).sum(dim=1) / torch.clamp((p * p).sum(dim=1), min=1e-5)).detach()
    coef = torch.clamp(coef, 0.001, 1000)
    return origin * coef.view([nb] + [1] * (origin.dim() - 1))


def ls_regress_diff_spec(diff, spec, imOrig, diffOrig, specOrig):
    """models.py:23-84: 2x2 per-sample normal equations with diffuse-only fallback, then a
    second scalar regression on clamp(d+s,0,1)."""
    nb, nc, nh, nw = diff.shape
    mask = (imOrig < 0.9).to(diff.dtype)
    d = (diff * mask).reshape(nb, -1)
    s = (spec * mask).reshape(nb, -1)
    im = (imOrig * mask).reshape(nb, -1)
    a11 = (d * d).sum(1)
    a22 = (s * s).sum(1)
    a12 = (d * s).sum(1)
    frac = a11 * a22 - a12 * a12
    b1 = (d * im).sum(1)
    b2 = (s * im).sum(1)
    coef1 = (b1 * a22 - b2 * a12) / torch.clamp(frac, min=1e-2)
    coef2 = (-b1 * a12 + a11 * b2) / torch.clamp(frac, min=1e-2)
    coef3 = torch.clamp(b1 / torch.clamp(a11, min=1e-5), 0.001, 1000)
    ind = ((frac / (nc * nh * nw)).detach() > 1e-2).to(diff.dtype)
    cD = torch.clamp(ind * coef1 + (1 - ind) * coef3, min=0, max=1000).view(nb, 1, 1, 1)
    cS = torch.clamp(ind * coef2, min=0, max=1000).view(nb, 1, 1, 1)
    dS = cD * diffOrig
    sS = cS * specOrig
    r = torch.clamp(dS + sS, 0, 1).reshape(nb, -1)
    io = imOrig.reshape(nb, -1)
    cI = ((r * io).sum(1) / torch.clamp((r * r).sum(1), min=1e-5)).detach()
    cI = torch.clamp(cI, 0.001, 1000).view(nb, 1, 1, 1)
    return cI * dS, cI * sS


def render_loss(diffusePred, specularPred, imSmall, segSmall):
    """wrapperBRDFLight.py:192-207: LSregressDiffSpec on detached stats, clamp, masked L2.

    Returns (renderErr, renderedImPred, numerator, pixelNum).
    """
    pixelNum = max(float(segSmall.sum()), 1e-5)
    dS, sS = ls_regress_diff_spec(diffusePred.detach(), specularPred.detach(), imSmall,
                                  diffusePred, specularPred)
    rend = torch.clamp(dS + sS, 0, 1)
    num = ((rend - imSmall) * (rend - imSmall) * segSmall.expand_as(imSmall)).sum()
    return num / pixelNum / 3.0, rend, num, pixelNum


def reconst_loss(envPred, envGT, segEnv, offset=1.0, envHeight=8, envWidth=16):
    """wrapperBRDFLight.py:179-188: LSregress on masked env, log-L2."""
    pixelNum = max(float(segEnv.sum()), 1e-5)
    m = segEnv.expand_as(envGT)
    scaled = ls_regress(envPred.detach() * m, envGT * m, envPred)
    dlog = torch.log(scaled + offset) - torch.log(envGT + offset)
    num = (dlog * dlog * m).sum()
    return num / pixelNum / 3.0 / envWidth / envHeight, scaled, num, pixelNum


def pred_to_shading(pred, envWidth=32, envHeight=16, SGNum=12, exact_warp=False):
    """utils.py:156-195: pred [1,7K,R,C] (NumPy) -> shading [3,R,C] in fp64, direction by direction.
    shading_c = max(sum_d cosEl_d sinEl_d sum_k tan(pi/2 .999 w_kc) exp(tan(pi/2 .999 lam_k)(a_k.w_d - 1)), 0)."""
    import numpy as np
    pred = np.asarray(pred)
    K = SGNum
    R, C = pred.shape[2], pred.shape[3]
    p = pred[0].astype(np.float64) if exact_warp else pred[0]   # exact_warp: the fp64 truth the parity protocol measures against
    axis = p[0:3 * K].reshape(K, 3, R, C).astype(np.float64)
    # the tan warp runs in pred's own dtype (float32 array x Python scalar stays float32, utils.py:178-186); everything
    # after it is promoted to float64 by the float64 direction table
    lamb = np.tan(np.pi / 2.0 * (p[3 * K:4 * K].reshape(K, 1, R, C) * 0.999)).astype(np.float64)
    weight = np.tan(np.pi / 2.0 * (p[4 * K:7 * K].reshape(K, 3, R, C) * 0.999)).astype(np.float64)
    out = np.zeros((3, R, C), dtype=np.float64)
    for i in range(envHeight):
        el = (i + 0.5) / envHeight * np.pi / 2.0
        for j in range(envWidth):
            az = ((j + 0.5) / envWidth - 0.5) * 2 * np.pi
            w = np.array([np.sin(el) * np.cos(az), np.sin(el) * np.sin(az), np.cos(el)])
            dot = (axis * w.reshape(1, 3, 1, 1)).sum(axis=1, keepdims=True)          # [K,1,R,C]
            out += (weight * np.exp(lamb * (dot - 1))).sum(axis=0) * (np.cos(el) * np.sin(el))
    return np.maximum(out, 0.0)


# --------------------------------------------------------------------------------------
# seeded synthetic inputs
# --------------------------------------------------------------------------------------


def synth_inputs(B, H, W, R, C, K, seed=1234, dtype=torch.float32, hi=1.0):
    """albedo~U(0,1); normal = s*normalize(N(0,1)^3+(0,0,1.5)), s~U(.7,.98); rough~U(-1,1);
    axis = normalize(N(0,1)^3); lambOrig, weightOrig ~ U(0,hi)."""
    g = torch.Generator().manual_seed(seed)
    albedo = torch.rand(B, 3, H, W, generator=g, dtype=torch.float64)
    nrm = torch.randn(B, 3, H, W, generator=g, dtype=torch.float64)
    nrm[:, 2] += 1.5
    nrm = nrm / nrm.norm(dim=1, keepdim=True)
    s = 0.7 + 0.28 * torch.rand(B, 1, H, W, generator=g, dtype=torch.float64)
    normal = nrm * s
    rough = torch.rand(B, 1, H, W, generator=g, dtype=torch.float64) * 2 - 1
    axis = torch.randn(B, K, 3, R, C, generator=g, dtype=torch.float64)
    axis = axis / axis.norm(dim=2, keepdim=True)
    lamb = torch.rand(B, K, R, C, generator=g, dtype=torch.float64) * hi
    weight = torch.rand(B, 3 * K, R, C, generator=g, dtype=torch.float64) * hi
    # round through fp32 so that fp32 and fp64 runs see identical inputs
    outs = [t.float().to(dtype).contiguous() for t in (albedo, normal, rough, axis, lamb, weight)]
    return outs
