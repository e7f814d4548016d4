#!/usr/bin/env python
"""Secondary measurements for DESIGN.md section 8 (NOT the driver's bench contract -- that is bench.py):
one JSON line per BASELINE.json config that is not the headline.

  python bench_configs.py [--which trainlight,unfused,bilateral,k24,pool,e2e_parts] [--iters 10]

  trainlight : config 2b / 3-geometry: maps 16x240x320 pooled to a 120x160 grid, output2env (env-map
               materialised for reconstErr) + fused forwardEnv, SG grads only (albedo detached, frozen BRDF nets)
  unfused    : sg_to_env + render_env fwd/bwd on a materialised env-map (B=4, 240x320), HBM-leaning kernels
  bilateral  : config 4: BilateralFunction fwd+bwd, 8x480x640, modes 0/2/4
  k24        : config 5 per-GPU shard: 4 x 480x640, 24 SG lobes, fused fwd+bwd
All timings: CUDA events on the current stream, median of --iters after 3 warm-ups.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def timed(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def rand_inputs(B, H, W, R, C, K, dev, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    alb = torch.rand(B, 3, H, W, device=dev, generator=g)
    n = torch.randn(B, 3, H, W, device=dev, generator=g)
    n[:, 2] += 1.5
    n = n / n.norm(dim=1, keepdim=True)
    if (H, W) == (R, C):
        n = n * (0.7 + 0.28 * torch.rand(B, 1, H, W, device=dev, generator=g))
    rg = torch.rand(B, 1, H, W, device=dev, generator=g) * 2 - 1
    ax = torch.randn(B, K, 3, R, C, device=dev, generator=g)
    ax = ax / ax.norm(dim=2, keepdim=True)
    lam = torch.rand(B, K, R, C, device=dev, generator=g)
    wt = torch.rand(B, 3 * K, R, C, device=dev, generator=g)
    return alb, n, rg, ax, lam, wt


def peak_hbm():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        return float(json.load(open(p))['hbm_gbs'])
    except Exception:
        return 6650.0


def run_trainlight(models, dev, iters):
    B, H, W, R, C, K = 16, 240, 320, 120, 160, 12
    alb, n, rg, ax, lam, wt = rand_inputs(B, H, W, R, C, K, dev, 1236)
    rl = models.renderingLayer(imWidth=C, imHeight=R, isCuda=True)
    o2e = models.output2env(K, isCuda=True)
    im = torch.rand(B, 3, R, C, device=dev)
    envgt = torch.rand(B, 3, R, C, 8, 16, device=dev)
    seg = (torch.rand(B, 1, R, C, device=dev) > 0.2).float()
    segenv = seg.unsqueeze(-1).unsqueeze(-1)

    def step():   # wrapperBRDFLight.py:177-207 with synthetic tensors
        a, l, w = [t.detach().requires_grad_(True) for t in (ax, lam, wt)]
        env, _, _, _ = o2e.output2env(a, l, w)
        m = segenv.expand_as(envgt)
        envs = models.LSregress(env.detach() * m, envgt * m, env)
        dl = torch.log(envs + 1) - torch.log(envgt + 1)
        rec = (dl * dl * m).sum() / (B * R * C) / 3.0 / 128
        d, s = rl.forwardEnv(alb, n, rg, env)
        dS, sS = models.LSregressDiffSpec(d.detach(), s.detach(), im, d, s)
        rend = torch.clamp(dS + sS, 0, 1)
        rerr = ((rend - im) * (rend - im) * seg.expand_as(im)).sum() / (B * R * C) / 3.0
        (rerr + 10 * rec).backward()
    ms = timed(step, iters)

    def step_fused():   # same losses; reconstruction term through output2env.reconstLoss, render term through forwardSG
        a, l, w = [t.detach().requires_grad_(True) for t in (ax, lam, wt)]
        rec, _ = o2e.reconstLoss(a, l, w, envgt, segenv)
        d, s = rl.forwardSG(alb, n, rg, a, l, w)
        dS, sS = models.LSregressDiffSpec(d.detach(), s.detach(), im, d, s)
        rend = torch.clamp(dS + sS, 0, 1)
        rerr = ((rend - im) * (rend - im) * seg.expand_as(im)).sum() / (B * R * C) / 3.0
        (rerr + 10 * rec).backward()
    ms_f = timed(step_fused, iters)

    def reconst_only():
        a, l, w = [t.detach().requires_grad_(True) for t in (ax, lam, wt)]
        rec, _ = o2e.reconstLoss(a, l, w, envgt, segenv)
        rec.backward()
    ms_rec = timed(reconst_only, iters)

    def render_only():
        a, l, w = [t.detach().requires_grad_(True) for t in (ax, lam, wt)]
        d, s = rl.forwardSG(alb, n, rg, a, l, w)
        torch.autograd.backward([d, s], [im, im])
    ms_r = timed(render_only, iters)
    npx = B * R * C
    return {'config': 'trainLight geometry: maps 16x240x320 -> grid 120x160, 12 SG', 'render_loss_step_ms': ms,
            'render_loss_step_envfree_ms': ms_f, 'reconst_loss_fwd_bwd_ms': ms_rec,
            'fused_render_fwd_bwd_sg_grads_only_ms': ms_r, 'render_Mpix_s': npx / ms_r / 1e3,
            'note': 'step = wrapperBRDFLight.py:177-207: output2env (env materialised) + masked LSregress + log-L2 + fused '
                    'forwardEnv + LSregressDiffSpec + render loss + backward; envfree = the same losses with '
                    'output2env.reconstLoss + forwardSG (the env-map is never written)'}


def run_trainlight_ddp(models, dev, iters):
    """config 3 at N ranks (launch under torch.distributed.run): every rank holds 16 images (weak scaling), runs the
    env-free loss block and normalises both losses by the batch-GLOBAL pixel counts through ONE 4-float all-reduce
    (sharding.sharded_losses, NCCL).  Device-timed, max over ranks, printed by rank 0."""
    import torch.distributed as dist
    from inverserenderingofindoorscene_b200 import sharding
    rank, world = dist.get_rank(), dist.get_world_size()
    B, H, W, R, C, K = 16, 240, 320, 120, 160, 12
    alb, n, rg, ax, lam, wt = rand_inputs(B, H, W, R, C, K, dev, 1236 + rank)
    rl = models.renderingLayer(imWidth=C, imHeight=R, isCuda=True)
    o2e = models.output2env(K, isCuda=True)
    g = torch.Generator(device=dev).manual_seed(77 + rank)
    im = torch.rand(B, 3, R, C, device=dev, generator=g)
    envgt = torch.rand(B, 3, R, C, 8, 16, device=dev, generator=g)
    seg = (torch.rand(B, 1, R, C, device=dev, generator=g) > 0.2).float()

    def step():
        a, l, w = [t.detach().requires_grad_(True) for t in (ax, lam, wt)]
        rec_num, rec_cnt, _ = o2e.reconstTerms(a, l, w, envgt, seg)
        d, s = rl.forwardSG(alb, n, rg, a, l, w)
        dS, sS = models.LSregressDiffSpec(d.detach(), s.detach(), im, d, s)
        rend = torch.clamp(dS + sS, 0, 1)
        r_num = ((rend - im) * (rend - im) * seg.expand_as(im)).sum()
        rerr, rec = sharding.sharded_losses(r_num, seg.sum(), rec_num, rec_cnt)
        (rerr + 10 * rec).backward()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    return {'config': 'config 3: trainLight render-loss step (env-free), 16 images per GPU, loss all-reduce over NCCL',
            'n_gpus': world, 'step_ms_max_over_ranks': ms, 'images_per_s': world * B / ms * 1e3, 'scaling': 'weak'}


def run_unfused(models, dev, iters, B=4):
    R, C, K = 240, 320, 12
    alb, n, rg, ax, lam, wt = rand_inputs(B, R, C, R, C, K, dev, 1237)
    rl = models.renderingLayer(imWidth=C, imHeight=R, isCuda=True)
    o2e = models.output2env(K, isCuda=True)
    npx = B * R * C
    res = {'config': 'unfused boundary kernels, %dx240x320, 12 SG (env-map %d MB)' % (B, B * 3 * R * C * 128 * 4 // 1000000)}
    a, l, w = [t.detach().requires_grad_(True) for t in (ax, lam, wt)]
    holder = {}

    def f1():
        holder['env'] = o2e.output2env(a, l, w)[0].materialize()   # the env-map is lazy: force the sg_to_env kernel
    res['sg_to_env_fwd_ms'] = timed(f1, iters)
    env = holder['env']
    genv = torch.rand_like(env)

    def f2():
        torch.autograd.grad(env, [a, l, w], genv, retain_graph=True)
    res['sg_to_env_bwd_ms'] = timed(f2, iters)
    envp = env.detach().clone().requires_grad_(True)
    leaves = [t.detach().requires_grad_(True) for t in (alb, n, rg)]

    def f3():
        holder['ds'] = rl.forwardEnv(leaves[0], leaves[1], leaves[2], envp)
    res['render_env_fwd_ms'] = timed(f3, iters)
    d, s = holder['ds']
    g1 = torch.rand_like(d)

    def f4():
        torch.autograd.grad([d, s], leaves + [envp], [g1, g1], retain_graph=True)
    res['render_env_bwd_ms'] = timed(f4, iters)
    pk = peak_hbm()
    byt = {'sg_to_env_fwd': (11 * K + 384) * 4, 'sg_to_env_bwd': (384 + 14 * K) * 4, 'render_env_fwd': (384 + 13) * 4,
           'render_env_bwd': (2 * 384 + 20) * 4}
    for k, bpp in byt.items():
        gbs = bpp * npx / (res[k + '_ms'] * 1e-3) / 1e9
        res[k + '_GBs'] = gbs
        res[k + '_hbm_frac'] = gbs / pk
    return res


def run_bilateral(BL, dev, iters):
    from oracle import bilateral_oracle as bo  # synthetic guide generator only (test infrastructure)
    B, H, W = 8, 480, 640
    guide = torch.from_numpy(bo.synth_guide(B, H, W, blur_sigma=8.0, seed=1238)).to(dev)
    out = {'config': 'BilateralFunction fwd+bwd, 8x480x640'}
    for mode, ch in ((0, 3), (2, 1), (4, 1)):
        sig, bsp = BL._PRESETS[mode]
        gp = torch.tensor([float(x) for x in sig], device=dev)
        bp = torch.tensor([float(x) for x in bsp], device=dev)
        pred = torch.rand(B, ch, H, W, device=dev)
        conf = 0.05 + 0.95 * torch.rand(B, 1, H, W, device=dev)
        gout = torch.randn(B, ch, H, W, device=dev)
        holder = {}

        def fwd():
            p = pred.detach().requires_grad_(True)
            c = conf.detach().requires_grad_(True)
            o, _ = BL.BilateralFunction.apply(guide, p, c, gp, bp)
            holder['x'] = (o, p, c)

        def fwdbwd():
            fwd()
            o, p, c = holder['x']
            torch.autograd.grad(o, [p, c], gout)
        f = timed(fwd, iters)
        fb = timed(fwdbwd, iters)
        out['mode%d' % mode] = {'ch': ch, 'fwd_ms': f, 'fwd_bwd_ms': fb, 'Mpix_s_fwd_bwd': B * H * W / fb / 1e3}
    return out


def run_k24(models, dev, iters):
    B, R, C, K = 4, 480, 640, 24
    alb, n, rg, ax, lam, wt = rand_inputs(B, R, C, R, C, K, dev, 1238)
    rl = models.renderingLayer(imWidth=C, imHeight=R, isCuda=True)
    g1 = torch.rand(B, 3, R, C, device=dev)
    holder = {}

    def fwd():
        leaves = [t.detach().requires_grad_(True) for t in (alb, n, rg, ax, lam, wt)]
        holder['x'] = (rl.forwardSG(*leaves), leaves)

    def fwdbwd():
        fwd()
        (d, s), leaves = holder['x']
        torch.autograd.backward([d, s], [g1, g1])
    f = timed(fwd, iters)
    fb = timed(fwdbwd, iters)
    return {'config': 'config 5 per-GPU shard: 4x480x640, 24 SG, fused fwd+bwd all grads', 'fwd_ms': f, 'fwd_bwd_ms': fb,
            'Mpix_s': B * R * C / fb / 1e3}


def run_torchcuda(dev, iters):
    """the oracle port (PyTorch op graph, direction-by-direction restatement of the reference) on the SAME GPU:
    what a torch-CUDA implementation of the layer costs next to the fused kernel (a comparator, not product)."""
    from oracle import render_oracle as ro
    B, R, C, K = 2, 240, 320, 12
    alb, n, rg, ax, lam, wt = rand_inputs(B, R, C, R, C, K, dev, 1239)
    g1 = torch.rand(B, 3, R, C, device=dev)

    def step():
        leaves = [t.detach().requires_grad_(True) for t in (alb, n, rg, ax, lam, wt)]
        d, s = ro.render_sg(*leaves)
        torch.autograd.backward([d, s], [g1, g1])
    ms = timed(step, max(3, iters // 3), warm=1)
    return {'config': 'oracle port as a torch-CUDA op graph, 2x240x320, 12 SG, fwd+bwd all grads', 'fwd_bwd_ms': ms,
            'Mpix_s': B * R * C / ms / 1e3}


def run_reference_torch_cuda(dev, iters=3, B=2):
    """BASELINE config 2's comparator: the UNMODIFIED reference render layer (baseline/_ref/models.py, isCuda=True) as a
    PyTorch op graph on the SAME GPU -- output2env.output2env + renderingLayer.forwardEnv + autograd backward, all six
    input gradients, 240x320, 12 SG.  Batch B (the reference materialises ~5 GB of temporaries per image)."""
    from oracle import ref_loader
    if not ref_loader.available():
        return {'unavailable': 'baseline/_ref not present'}
    ref = ref_loader.load('models')
    R, C, K = 240, 320, 12
    torch.cuda.set_device(dev)
    alb, n, rg, ax, lam, wt = rand_inputs(B, R, C, R, C, K, dev, 1239)
    g1 = torch.rand(B, 3, R, C, device=dev)
    rl = ref.renderingLayer(imWidth=C, imHeight=R, isCuda=True)
    o2e = ref.output2env(K, isCuda=True)

    def step():
        leaves = [t.detach().requires_grad_(True) for t in (alb, n, rg, ax, lam, wt)]
        env, _, _, _ = o2e.output2env(leaves[3], leaves[4], leaves[5])
        d, s = rl.forwardEnv(leaves[0], leaves[1], leaves[2], env)
        torch.autograd.backward([d, s], [g1, g1])
    ms = timed(step, iters, warm=1)
    peak_gb = torch.cuda.max_memory_allocated(dev) / 2**30
    torch.cuda.empty_cache()
    return {'what': 'reference models.py (isCuda=True) torch-CUDA op graph on this GPU, %dx240x320, 12 SG, fwd+bwd all grads' % B,
            'fwd_bwd_ms': ms, 'Mpix_s': B * R * C / ms / 1e3, 'peak_mem_GiB': peak_gb}


def run_bilateral_mode0(BL, dev, iters=5):
    """config 4 headline: BilateralFunction fwd+bwd, 8x480x640, mode 0 (3 channels)."""
    import ctypes as C
    from inverserenderingofindoorscene_b200 import _lib
    from oracle import bilateral_oracle as bo  # synthetic guide generator only (test infrastructure)
    B, H, W, ch = 8, 480, 640, 3
    guide = torch.from_numpy(bo.synth_guide(B, H, W, blur_sigma=8.0, seed=1238)).to(dev)
    sig, bsp = BL._PRESETS[0]
    gp = torch.tensor([float(x) for x in sig], device=dev)
    bp = torch.tensor([float(x) for x in bsp], device=dev)
    g = torch.Generator(device=dev).manual_seed(5)
    pred = torch.rand(B, ch, H, W, device=dev, generator=g)
    conf = 0.05 + 0.95 * torch.rand(B, 1, H, W, device=dev, generator=g)
    gout = torch.randn(B, ch, H, W, device=dev, generator=g)
    holder = {}

    def fwd():
        p = pred.detach().requires_grad_(True)
        c = conf.detach().requires_grad_(True)
        o, _ = BL.BilateralFunction.apply(guide, p, c, gp, bp)
        holder['x'] = (o, p, c)

    def fwdbwd():
        fwd()
        o, p, c = holder['x']
        torch.autograd.grad(o, [p, c], gout)
    lib = _lib.lib()
    l0 = lib.irb_launch_count()
    fwdbwd()
    launches = lib.irb_launch_count() - l0
    f = timed(fwd, iters)
    fb = timed(fwdbwd, iters)
    # vertex / link counts of the grids (one extra build with the host read-back of V; E from the neighbour table)
    nb = lib.irb_bg_workspace_bytes(B, H, W, ch)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    nv = (C.c_int * B)()
    _lib.check(lib.irb_bg_build(_lib.ptr(guide), 255.0, B, H, W, float(sig[0]), float(sig[1]), float(sig[2]), _lib.ptr(ws), nb, nv,
                                _lib.stream_ptr(dev)), 'irb_bg_build')
    V = [int(x) for x in nv]
    P = H * W
    # streaming model of SURVEY 8d with THIS build's storage (fp64 vectors): per PCG iteration and vertex-channel
    # 43 B (p-update) + 115 B (mat-vec) + 56 B (x/r update); 12 iterations forward + 12 adjoint; + compulsory pixel traffic
    vsum = float(sum(V))
    model_bytes = 2 * 12 * ch * vsum * (43 + 115 + 56) + B * P * 100.0
    return {'what': 'BilateralFunction fwd+bwd, 8x480x640, mode 0 (3 ch, sigma 8/2/7, lam 200, 12 PCG iterations)',
            'fwd_ms': f, 'fwd_bwd_ms': fb, 'Mpix_s_fwd_bwd': B * P / fb / 1e3, 'launches_fwd_bwd': int(launches),
            'V_per_image': V, 'V_over_P': vsum / (B * P),
            'streaming_model_GB': model_bytes / 1e9, 'streaming_model_GBs': model_bytes / (fb * 1e-3) / 1e9,
            'hbm_frac_of_model': model_bytes / (fb * 1e-3) / 1e9 / peak_hbm()}


def run_ctor_cost(models):
    """testReal.py:341-346 builds a renderingLayer per image with that image's (H, W, fov): first construction of a geometry
    (host arithmetic + 5 H2D copies) against a repeat (table cache hit)"""
    import time
    res = {}
    for tag, (w, h, fov) in (('160x120_fov57', (160, 120, 57)), ('213x120_fov57', (213, 120, 57)), ('160x90_fov42.75', (160, 90, 42.75))):
        models._LAYER_TABLES.clear()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        models.renderingLayer(imWidth=w, imHeight=h, fov=fov, isCuda=True)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        models.renderingLayer(imWidth=w, imHeight=h, fov=fov, isCuda=True)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        res[tag] = {'first_ms': (t1 - t0) * 1e3, 'cached_ms': (t2 - t1) * 1e3}
    return res


def graphed_step_ms(step, timed_ranks, dev, world):
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    torch.cuda.synchronize()
    ms = timed_ranks(g.replay, 20, 3, dev, world)
    del g
    return {'ms_per_step': ms, 'note': 'torch.cuda.CUDAGraph capture of the whole step (forward, loss, all-reduce, backward), replay timed'}


def extras_for_bench(models, sharding, dev, rank, world, timed_ranks, make_render_step):
    """secondary BASELINE configs for the bench line (`extras`): device-timed, max over ranks, a handful of steps each"""
    from inverserenderingofindoorscene_b200 import BilateralLayer as BL
    out = {}
    # config 5: 480x640, 24 SG, 4 images per GPU (B = 32 over 8 GPUs), reference call pattern + loss all-reduce
    step5 = make_render_step(models, sharding, dev, 4, 480, 640, 480, 640, 24, 1300 + rank)
    ms5 = timed_ranks(step5, 5, 3, dev, world)
    out['config5'] = {'what': 'config 5: 480x640, 24 SG, 4 images per GPU (global batch %d), fwd+bwd all grads + loss all-reduce'
                              % (4 * world), 'ms_per_step': ms5, 'Mpix_s': world * 4 * 480 * 640 / ms5 / 1e3}
    del step5
    # config 3 geometry (trainLight: maps 240x320 pooled to a 120x160 grid, SG grads only): render-loss block, strong
    # (global batch 16) and weak (16 per GPU)
    for name, b in (('config3_strong', max(1, 16 // world)), ('config3_weak', 16)):
        if world == 1 and name == 'config3_weak':
            continue
        st = make_render_step(models, sharding, dev, b, 240, 320, 120, 160, 12, 1400 + rank, all_grads=False)
        ms = timed_ranks(st, 10, 3, dev, world)
        out[name] = {'what': 'config 3 geometry: maps 240x320 -> grid 120x160, 12 SG, %d images per GPU (global batch %d), '
                             'render-loss step (output2env -> forwardEnv -> render loss -> all-reduce -> backward to the SG '
                             'parameters); the LightNet CNNs are out of scope' % (b, b * world),
                     'ms_per_step': ms, 'images_per_s': b * world / ms * 1e3}
        if name == 'config3_strong':
            # small per-GPU batches are launch / host bound (6 kernels + ~25 tiny torch ops per step): the same step captured
            # once in a CUDA graph (the library neither allocates nor synchronises, so its launches are capturable) and replayed
            try:
                out[name]['cuda_graph'] = graphed_step_ms(st, timed_ranks, dev, world)
            except Exception as ex:
                out[name]['cuda_graph'] = {'error': repr(ex)[:200]}
        del st
    torch.cuda.empty_cache()
    if world == 1:
        try:
            r = run_trainlight(models, dev, 5)
            out['config3_wrapper_pattern'] = {
                'what': 'the literal loss block of wrapperBRDFLight.py:177-207 on 16 images (120x160 grid): output2env (env-map MATERIALISED '
                        'because the wrapper reads it) + LSregress + torch log-L2 + forwardEnv (fused) + LSregressDiffSpec + torch render '
                        'loss + backward; env-free = the same losses through output2env.reconstLoss + forwardSG',
                'ms_per_step': r['render_loss_step_ms'], 'envfree_ms_per_step': r['render_loss_step_envfree_ms'],
                'reconst_loss_fwd_bwd_ms': r['reconst_loss_fwd_bwd_ms']}
        except Exception as ex:
            out['config3_wrapper_pattern'] = {'error': repr(ex)[:200]}
        torch.cuda.empty_cache()
        out['layer_construction'] = run_ctor_cost(models)
        out['config4_bilateral'] = run_bilateral_mode0(BL, dev)
        torch.cuda.empty_cache()
        out['reference_torch_cuda'] = run_reference_torch_cuda(dev)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--which', default='trainlight,unfused,bilateral,k24,torchcuda')
    ap.add_argument('--unfused-batch', type=int, default=4, help='images in the unfused-kernel config (4: 472 MB env-map; 16: 1.9 GB)')
    ap.add_argument('--iters', type=int, default=10)
    args = ap.parse_args()
    from inverserenderingofindoorscene_b200 import BilateralLayer as BL
    from inverserenderingofindoorscene_b200 import _lib, models
    _lib.lib()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    if world > 1 or 'trainlight_ddp' in args.which:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29533')
        os.environ.setdefault('RANK', '0')
        os.environ.setdefault('WORLD_SIZE', '1')
        dist.init_process_group('nccl', device_id=dev)
        if world > 1:   # only the sharded config makes sense on several ranks
            r = run_trainlight_ddp(models, dev, args.iters)
            if dist.get_rank() == 0:
                print(json.dumps(r))
            dist.destroy_process_group()
            return
    for w in args.which.split(','):
        if w == 'trainlight':
            r = run_trainlight(models, dev, args.iters)
        elif w == 'trainlight_ddp':
            r = run_trainlight_ddp(models, dev, args.iters)
        elif w == 'unfused':
            r = run_unfused(models, dev, args.iters, args.unfused_batch)
        elif w == 'bilateral':
            r = run_bilateral(BL, dev, args.iters)
        elif w == 'k24':
            r = run_k24(models, dev, args.iters)
        elif w == 'torchcuda':
            r = run_torchcuda(dev, args.iters)
        else:
            continue
        print(json.dumps(r))
        sys.stdout.flush()
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
