#!/usr/bin/env python
"""bench.py -- headline benchmark: Mpixels/s of the render layer forward+backward
(BASELINE.json config 2: batch 16 x 240x320, 12 SG lobes, fused SG shading, all input grads).

  python bench.py --gpus N --steps K --warmup W              (ours; N>1 under torchrun)
  python bench.py --impl reference --gpus N --steps K ...    (CPU arm: oracle port on host cores)

One "step" = one forward + one backward of the fused render layer over one synthetic batch.
`value`  : device-timed (CUDA events) throughput with inputs resident in HBM, max over ranks.
`e2e`    : same metric through the public API (models.renderingLayer.forwardSG + autograd) with
           pinned HOST buffers, H2D of the step's inputs and D2H of its result inside the timed region.
`roofline`: dominant kernel (render_sg_bwd) algorithmic bytes / its CUDA-event duration vs the
           measured HBM copy peak (MEASURED_PEAKS.json); the kernel is FP32/MUFU-bound, so an
           `fp32` sub-object reports the instruction-model FLOP rate beside it.
Multi-GPU: the batch dimension shards over ranks (weak scaling, 16 images per GPU); the only
exchange is the 4-float loss all-reduce (wrapperBRDFLight.py:192,205), done once per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

B_PER_GPU, H, W, K = 16, 240, 320, 12
METRIC = 'Mpixels/s render-layer fwd+bwd (240x320, 12 SG)'
ENV_H, ENV_W = 8, 16


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--profile', default=None, metavar='DIR',
                    help='capture `ncu --set full` of the fused kernels of THIS command (1 GPU) and write DIR/ncu_metrics.json + '
                         'ncu_summary.txt (e.g. profiles/r02) -- the file the roofline traffic / pipe fields are read from')
    ap.add_argument('--no-extras', action='store_true', help='skip the secondary configs (bilateral, config 3/5, torch-CUDA reference)')
    ap.add_argument('--cpu-baseline-rows', type=int, default=0, help='rows of the 240x320 frame for the CPU sample (0=auto)')
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs, burst copy)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


# -------------------------------------------------------------------------------------------
# synthetic inputs, generated on the host once, seeded
# -------------------------------------------------------------------------------------------

def synth(B, Hh, Ww, Kk, seed, device=None, pinned=False):
    g = torch.Generator().manual_seed(seed)
    albedo = torch.rand(B, 3, Hh, Ww, generator=g)
    n = torch.randn(B, 3, Hh, Ww, generator=g)
    n[:, 2] += 1.5
    n = n / n.norm(dim=1, keepdim=True) * (0.7 + 0.28 * torch.rand(B, 1, Hh, Ww, generator=g))
    rough = torch.rand(B, 1, Hh, Ww, generator=g) * 2 - 1
    axis = torch.randn(B, Kk, 3, Hh, Ww, generator=g)
    axis = axis / axis.norm(dim=2, keepdim=True)
    lamb = torch.rand(B, Kk, Hh, Ww, generator=g)
    weight = torch.rand(B, 3 * Kk, Hh, Ww, generator=g)
    gD = torch.rand(B, 3, Hh, Ww, generator=g)
    gS = torch.rand(B, 3, Hh, Ww, generator=g)
    ts = [albedo, n.contiguous(), rough, axis.contiguous(), lamb, weight, gD, gS]
    if pinned:
        ts = [t.pin_memory() for t in ts]
    if device is not None:
        ts = [t.to(device) for t in ts]
    return ts


# -------------------------------------------------------------------------------------------
# CPU arm: the reference's OWN render layer (baseline/_ref = unmodified copy of the reference's *.py, imported by
# oracle/ref_loader.py; kind "reference") on host cores -- or, if that copy is absent, the oracle port (kind "port").
# One step = output2env.output2env + renderingLayer.forwardEnv + autograd backward (all six input gradients) over a
# ROWS x 320 crop of one 240x320 frame (pixels are independent; the crop keeps the full frame's view vectors).
# Thread count and sample are FIXED (not probed) so that every CPU figure of a round is taken the same way.
# -------------------------------------------------------------------------------------------

CPU_THREADS = 32          # torch CPU elementwise ops on these tensors stop scaling around 16-32 threads (profiles/r01)
CPU_ROWS = 240            # cpu_baseline sample: one full 240 x 320 frame per step (= BASELINE config 1; ~4 s per step)


def cpu_threads():
    n = min(CPU_THREADS, os.cpu_count() or 1)
    torch.set_num_threads(n)
    return n


def cpu_step_fn(rows):
    """returns (step, kind)"""
    from oracle import ref_loader
    ins = synth(1, rows, W, K, seed=1234 + 1)
    leaves = [t.clone().requires_grad_(True) for t in ins[:6]]
    gD, gS = ins[6], ins[7]
    from oracle import render_oracle as ro
    v_full = torch.from_numpy(ro.camera_rays(H, W))[:, :rows].contiguous()   # [3,rows,W] of the full 240x320 frame
    if ref_loader.available():
        ref = ref_loader.load('models')
        rl = ref.renderingLayer(imWidth=W, imHeight=rows, isCuda=False)
        rl.v = v_full.unsqueeze(0)                       # public attribute (models.py:432): crop of the full frame's rays
        o2e = ref.output2env(K, isCuda=False)
        kind = 'reference'

        def render(alb, nrm, rg, ax, lam, wt):
            env, _, _, _ = o2e.output2env(ax, lam, wt)
            return rl.forwardEnv(alb, nrm, rg, env)
    else:
        kind = 'port'

        def render(alb, nrm, rg, ax, lam, wt):
            env, _, _, _ = ro.output2env(ax, lam, wt)
            return ro.render_env(alb, nrm, rg, env, view=v_full)

    def step():
        d, s = render(*leaves)
        torch.autograd.grad((d * gD).sum() + (s * gS).sum(), leaves)
    return step, kind


def _cpu_sample_text(kind, rows, threads):
    what = ('the reference itself: baseline/_ref/models.py output2env.output2env + renderingLayer.forwardEnv + autograd'
            if kind == 'reference' else 'oracle port (oracle/render_oracle.py restatement of the same op graph)')
    return ('%s, fp32 fwd+bwd all grads, %d x %d pixels of one 240x320 frame per step, 12 SG, torch %s CPU, %d threads '
            '(fixed; host has %d cores)' % (what, rows, W, torch.__version__, threads, os.cpu_count() or 1))


def cpu_baseline(rows=0):
    """bounded sample of config 2's workload on the host: best of 2 steps after 1 warm-up"""
    threads = cpu_threads()
    rows = rows if rows > 0 else CPU_ROWS
    step, kind = cpu_step_fn(rows)
    step()
    ts = []
    for _ in range(2):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    t = min(ts)
    return {'value': rows * W / t / 1e6, 'unit': 'Mpix/s', 'cores': threads, 'kind': kind,
            'sample': _cpu_sample_text(kind, rows, threads) + ', best of 2 after 1 warm-up'}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = cpu_threads()
    # per-step sample sized so that (steps + warmup) steps end within ~100 s, from one calibration step at CPU_ROWS
    # per-step sample: the full frame (the same sample cpu_baseline times) whenever (steps + warmup) full-frame steps end
    # within ~150 s, else proportionally fewer rows (calibrated with one 48-row step)
    nstep = max(1, args.steps + args.warmup)
    step, kind = cpu_step_fn(48)
    step()
    t0 = time.perf_counter()
    step()
    per_row = (time.perf_counter() - t0) / 48
    rows = int(max(8, min(H, 150.0 / nstep / max(per_row, 1e-6))))
    if rows > 200:
        rows = H
    step, kind = cpu_step_fn(rows)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = rows * W * args.steps / dt / 1e6
    sample = _cpu_sample_text(kind, rows, threads)
    emit(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': val,
        'unit': 'Mpix/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'config2: SG render layer fwd+bwd (output2env.output2env -> renderingLayer.forwardEnv), '
                               '240x320, 12 SG, all input grads; CPU sample: %d rows of one frame per step (per-pixel '
                               'rate: pixels are independent)' % rows},
        'cpu_baseline': {'value': val, 'unit': 'Mpix/s', 'cores': threads, 'kind': kind, 'sample': sample},
        'e2e': {'value': val, 'unit': 'Mpix/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0}))


# -------------------------------------------------------------------------------------------
# clocks sampler
# -------------------------------------------------------------------------------------------

class Clocks(threading.Thread):
    """samples SM clock + throttle reasons DURING the timed region: NVML in-process (sub-millisecond per sample),
    nvidia-smi as fallback"""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []     # (sm_mhz, max_mhz, [reason names])
        self.stop_flag = False
        self.source = 'nvml'
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # honour CUDA_VISIBLE_DEVICES-style remapping: take the handle by the torch device's PCI bus id
            import torch as _t
            prop = _t.cuda.get_device_properties(index)
            bus = getattr(prop, 'pci_bus_id', None)
            self.h = None
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    hh = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if pynvml.nvmlDeviceGetPciInfo(hh).bus == bus:
                        self.h = hh
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.nv = None
            self.source = 'nvidia-smi'

    def _sample_nvml(self):
        nv = self.nv
        sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
        mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        names = []
        for bit, name in ((0x8, 'hw_slowdown'), (0x40, 'hw_thermal_slowdown'), (0x20, 'sw_thermal_slowdown'),
                          (0x4, 'sw_power_cap')):
            if r & bit:
                names.append(name)
        self.samples.append((sm, mx, names))

    def _sample_smi(self):
        out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                              '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
        f = [x.strip() for x in out.strip().split(',')]
        if len(f) >= 6 and f[0].isdigit():
            names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
            self.samples.append((int(f[0]), int(f[1]), [n for i, n in enumerate(names) if f[2 + i].lower().startswith('active')]))

    def run(self):
        while not self.stop_flag:
            try:
                if self.nv is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            time.sleep(0.005)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        sm = sorted(s[0] for s in self.samples)
        reasons = sorted(set(n for s in self.samples for n in s[2]))
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': max(s[1] for s in self.samples), 'reasons': reasons,
                'samples': len(self.samples), 'source': self.source}


# -------------------------------------------------------------------------------------------
# ours
# -------------------------------------------------------------------------------------------

def ncu_metrics():
    """per-kernel counters of the last `python bench.py --profile` run on a B200 (profiles/<round>/ncu_metrics.json,
    written by tools/ncu_profile.py from an `ncu --set full` capture of THIS bench command): DRAM traffic per launch,
    pipe utilisations, registers.  Returns {} when no capture is committed."""
    best = {}
    pdir = os.path.join(ROOT, 'profiles')
    if os.path.isdir(pdir):
        for rnd in sorted(os.listdir(pdir)):
            f = os.path.join(pdir, rnd, 'ncu_metrics.json')
            if os.path.exists(f):
                try:
                    best = json.load(open(f))
                    best['_file'] = os.path.relpath(f, ROOT)
                except Exception:
                    pass
    return best


def bind_to_gpu_numa(index):
    """run this rank (and therefore its pinned allocations: cudaHostAlloc pages are first-touched by the calling thread) on
    the CPUs of the NUMA node the GPU hangs off -- at N = 8 the pinned host buffers of all ranks otherwise land on whichever
    node the launcher started on and every H2D copy crosses the socket interconnect"""
    try:
        prop = torch.cuda.get_device_properties(index)
        bus = '%04x:%02x:%02x.0' % (getattr(prop, 'pci_domain_id', 0), prop.pci_bus_id, prop.pci_device_id)
        base = '/sys/bus/pci/devices/' + bus
        cpulist = open(base + '/local_cpulist').read().strip()
        node = open(base + '/numa_node').read().strip()
        cpus = set()
        for part in cpulist.split(','):
            if '-' in part:
                a, b = part.split('-')
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {'numa_node': int(node), 'cpus': len(cpus), 'pci': bus}
    except Exception as ex:
        return {'numa_node': None, 'error': repr(ex)[:80]}
    return {'numa_node': None}


def max_over_ranks(ms, dev, world):
    import torch.distributed as dist
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def timed_ranks(step, steps, warmup, dev, world):
    """W warm-ups, then exactly `steps` steps between barrier+synchronize pairs, CUDA events, max over ranks (ms/step)"""
    import torch.distributed as dist

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
    for _ in range(warmup):
        step()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    sync_all()
    return max_over_ranks(e0.elapsed_time(e1) / steps, dev, world)


def make_render_step(irb_models, sharding, dev, B, Hm, Wm, R, C, Kk, seed, all_grads=True):
    """One render-loss step through the reference call pattern (wrapperBRDFLight.py:177,194-207):
    output2env.output2env -> renderingLayer.forwardEnv -> render loss (fused LSregressDiffSpec + clamp + masked L2)
    -> 4-float loss all-reduce (NCCL when world > 1) -> backward."""
    g = torch.Generator(device=dev).manual_seed(seed)
    alb = torch.rand(B, 3, Hm, Wm, device=dev, generator=g)
    n = torch.randn(B, 3, Hm, Wm, device=dev, generator=g)
    n[:, 2] += 1.5
    n = n / n.norm(dim=1, keepdim=True)
    if (Hm, Wm) == (R, C):
        n = n * (0.7 + 0.28 * torch.rand(B, 1, Hm, Wm, device=dev, generator=g))
    rg = torch.rand(B, 1, Hm, Wm, device=dev, generator=g) * 2 - 1
    ax = torch.randn(B, Kk, 3, R, C, device=dev, generator=g)
    ax = ax / ax.norm(dim=2, keepdim=True)
    lam = torch.rand(B, Kk, R, C, device=dev, generator=g)
    wt = torch.rand(B, 3 * Kk, R, C, device=dev, generator=g)
    im = torch.rand(B, 3, R, C, device=dev, generator=g)
    seg = (torch.rand(B, 1, R, C, device=dev, generator=g) > 0.2).float()
    layer = irb_models.renderingLayer(imWidth=C, imHeight=R, isCuda=True)
    o2e = irb_models.output2env(Kk, isCuda=True)
    maps = (alb, n.contiguous(), rg)

    def step():
        m = [t.detach().requires_grad_(all_grads) for t in maps]
        sg = [t.detach().requires_grad_(True) for t in (ax, lam, wt)]
        env, _, _, _ = o2e.output2env(*sg)
        d, s = layer.forwardEnv(m[0], m[1], m[2], env)
        _, num, cnt, _ = irb_models.renderLossTerms(d, s, im, seg)
        rerr, _ = sharding.sharded_losses(num, cnt, None, None)
        rerr.backward()
        return rerr
    return step


def run_ours(args):
    import torch.distributed as dist
    from inverserenderingofindoorscene_b200 import models as irb_models
    from inverserenderingofindoorscene_b200 import _lib, sharding

    lib = _lib.lib()  # fail loudly if the CUDA library is missing
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device (there is no CPU path for --impl ours)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    numa = bind_to_gpu_numa(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    B = B_PER_GPU
    host = synth(B, H, W, K, seed=1234 + 2 + rank, pinned=True)
    gi = torch.Generator().manual_seed(99 + rank)
    im_host = torch.rand(B, 3, H, W, generator=gi).pin_memory()
    seg_host = (torch.rand(B, 1, H, W, generator=gi) > 0.2).float().pin_memory()
    devt = [t.to(dev) for t in host[:6]]
    im_d, seg_d = im_host.to(dev), seg_host.to(dev)
    layer = irb_models.renderingLayer(imWidth=W, imHeight=H, isCuda=True)
    o2e = irb_models.output2env(K, isCuda=True)
    npx = B * H * W
    warm = max(3, args.warmup)

    # The step, through the REFERENCE call pattern (wrapperBRDFLight.py:177,194-207): output2env.output2env (lazy env-map,
    # never written because forwardEnv is its only consumer) -> renderingLayer.forwardEnv (fused kernel) -> render loss
    # (irb_render_loss: LSregressDiffSpec + clamp + masked L2, device-side pixelNum) -> the path's ONE exchange, the
    # 4-float loss all-reduce (NCCL at N > 1) -> backward (loss seed kernel + fused render backward, all six input grads).
    def step_device(ins, im, seg, ev=None):
        leaves = [t.detach().requires_grad_(True) for t in ins]
        if ev:
            ev[0].record()
        env, _, _, _ = o2e.output2env(leaves[3], leaves[4], leaves[5])
        d, s = layer.forwardEnv(leaves[0], leaves[1], leaves[2], env)
        if ev:
            ev[1].record()
        rend, num, cnt, _ = irb_models.renderLossTerms(d, s, im, seg)
        rerr, _ = sharding.sharded_losses(num, cnt, None, None)
        rerr.backward()
        return rend, rerr, leaves

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- value: device-resident inputs -------------------------------------------------------
    for _ in range(warm):
        step_device(devt, im_d, seg_d)
    sync_all()
    clocks = Clocks(local)
    clocks.start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(args.steps)]
    sink = []
    irb_models._EVENT_SINK = sink          # the fused backward kernel brackets itself with CUDA events on its stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    l0 = lib.irb_launch_count()
    e0.record()
    for i in range(args.steps):
        step_device(devt, im_d, seg_d, evs[i])
    e1.record()
    sync_all()
    launches = lib.irb_launch_count() - l0
    irb_models._EVENT_SINK = None
    clocks.stop_flag = True
    ms_step = max_over_ranks(e0.elapsed_time(e1) / args.steps, dev, world)
    fwd_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    bwd_evs = [x for x in sink if x[0] == 'render_sg_bwd']
    bwd_ms = sum(x[1].elapsed_time(x[2]) for x in bwd_evs) / max(1, len(bwd_evs))
    value = world * npx / (ms_step * 1e-3) / 1e6

    # ---- e2e: pinned host buffers -> public API -> loss terms + rendered image back to host ---------
    # Every step copies ITS inputs host->device and reads ITS results back; like any production feeder the copies
    # of step i+1 run on a second stream while step i computes (double-buffered device inputs).
    out_host = torch.empty(B, 3, H, W).pin_memory()
    loss_host = torch.empty(1).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    srcs = list(host[:6]) + [im_host, seg_host]
    bufs = [[torch.empty_like(h, device=dev) for h in srcs] for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    freed = [torch.cuda.Event(), torch.cuda.Event()]

    def upload(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[slot])          # the compute stream is done with this slot
            for dst, src in zip(bufs[slot], srcs):
                dst.copy_(src, non_blocking=True)
            ready[slot].record(copy_stream)

    def step_e2e(slot, prefetch):
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(ready[slot])
        if prefetch:
            upload(slot ^ 1)
        rend, rerr, _ = step_device(bufs[slot][:6], bufs[slot][6], bufs[slot][7])
        freed[slot].record(cur)
        out_host.copy_(rend, non_blocking=True)
        loss_host.copy_(rerr.detach().reshape(1), non_blocking=True)
        cur.synchronize()                  # the step's results are on the host
        return float(loss_host[0])

    def run_e2e(n):
        for e in freed:
            e.record(torch.cuda.current_stream(dev))
        upload(0)
        for i in range(n):
            step_e2e(i & 1, i + 1 < n)

    run_e2e(3)
    sync_all()
    n_e2e = max(4, min(args.steps, 12))
    e0.record()
    run_e2e(n_e2e)
    e1.record()
    sync_all()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1) / n_e2e, dev, world)
    h2d = sum(h.numel() * 4 for h in srcs)
    d2h = out_host.numel() * 4 + loss_host.numel() * 4

    # ---- the other BASELINE configs that run on this many GPUs (extras; not the headline) --------------
    extras = {}
    if not args.no_extras:
        import bench_configs as bc
        try:
            extras = bc.extras_for_bench(irb_models, sharding, dev, rank, world, timed_ranks, make_render_step)
        except Exception as ex:   # extras never take the headline down
            extras = {'error': repr(ex)}

    if rank == 0:
        peak, peak_src = peaks()
        # algorithmic bytes per pixel: the forward reads 7K+7 floats and writes 6 + the 2K-float side tensor of
        # per-lobe hemisphere sums; the backward reads 7K+13 + that side tensor and writes 7K+7
        bwd_bytes = (20 + 14 * K + 2 * K) * 4 * npx
        fwd_bytes = (13 + 7 * K + 2 * K) * 4 * npx
        ach = bwd_bytes / (bwd_ms * 1e-3) / 1e9
        clk = clocks.summary()
        nm = ncu_metrics()
        kb = nm.get('render_sg_bwd', {})
        kf = nm.get('render_sg_fwd', {})
        line = {
            'metric': METRIC, 'value': value, 'unit': 'Mpix/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': warm, 'ms_per_step': ms_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'config2: SG render layer fwd+bwd through the reference call pattern '
                                   'models.output2env.output2env -> models.renderingLayer.forwardEnv (lazy env-map, fused '
                                   'kernel) + fused render loss + 4-float loss all-reduce + backward, batch 16 per GPU x '
                                   '240x320, 12 SG, all six input grads',
                       'global_batch': B * world, 'grid': [H, W], 'sg_lobes': K, 'parallelism': 'batch-shard x%d' % world,
                       'collective': 'one 4-float all-reduce per step inside the timed region (NCCL)' if world > 1
                                     else 'none at N=1 (same code path, world size 1)',
                       'l2': 'inputs per step (447 MB/GPU) exceed the 126 MB L2; no explicit flush'},
            'kernel_ms': {'render_sg_fwd': fwd_ms, 'render_sg_bwd': bwd_ms,
                          'loss_allreduce_seed_and_host_gaps': ms_step - fwd_ms - bwd_ms},
            'roofline': {'bound': 'hbm', 'binding_resource': 'fp32 FMA pipe (compute-bound kernel, ~70 FLOP/B; DESIGN.md 4.1)',
                         'kernel': 'render_sg_group_kernel<1,...> (fused backward, lobe-split)', 'achieved': ach,
                         'peak': peak, 'unit': 'GB/s', 'frac': ach / peak,
                         'traffic': kb.get('dram_bytes'), 'traffic_source': nm.get('_file'),
                         'peak_source': peak_src,
                         'algorithmic_bytes_per_launch': bwd_bytes,
                         'fwd': {'achieved': fwd_bytes / (fwd_ms * 1e-3) / 1e9, 'algorithmic_bytes_per_launch': fwd_bytes,
                                 'traffic': kf.get('dram_bytes')},
                         'pipes_ncu': {'source': nm.get('_file'), 'note': 'counters of the committed ncu --set full capture of this '
                                       'bench command (python bench.py --profile); not re-measured in this run',
                                       'bwd': {k: kb.get(k) for k in ('fma_pipe_pct', 'xu_pipe_pct', 'issue_active_pct',
                                                                      'registers', 'duration_us', 'warps_per_sm')},
                                       'fwd': {k: kf.get(k) for k in ('fma_pipe_pct', 'xu_pipe_pct', 'issue_active_pct',
                                                                      'registers', 'duration_us', 'warps_per_sm')}}},
            'e2e': {'value': world * npx / (e2e_ms * 1e-3) / 1e6, 'unit': 'Mpix/s', 'ms_per_step': e2e_ms,
                    'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'steps': n_e2e,
                    'h2d_GBs_per_rank': h2d / (e2e_ms * 1e-3) / 1e9, 'host_binding_rank0': numa},
            'gpu_launches': int(launches),   # irb_launch_count() delta over the timed region (this rank)
            'gpu_launches_per_step': launches / args.steps,
            'clocks': clk,
            'extras': extras,
        }
        if not args.no_cpu_baseline and world == 1:
            line['cpu_baseline'] = cpu_baseline(rows=args.cpu_baseline_rows)
        elif world > 1:
            line['cpu_baseline'] = {'value': None, 'unit': 'Mpix/s', 'cores': os.cpu_count(), 'kind': 'reference',
                                    'sample': 'reported at N=1 only'}
        emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_profile(args):
    """ncu replays every captured kernel ~40x: never a timing run, never multi-rank"""
    tool = os.path.join(ROOT, 'tools', 'ncu_profile.py')
    scratch = os.path.join(ROOT, 'gpurun_out', 'ncu_bench')
    cmd = [sys.executable, tool, 'capture', '--out', scratch, '--kernels', 'render_sg_group_kernel|render_loss|lsreg', '--skip',
           '20', '--count', '6', '--', sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '2', '--warmup', '3',
           '--no-extras', '--no-cpu-baseline']
    rc = subprocess.run(cmd).returncode
    if rc == 0:
        rc = subprocess.run([sys.executable, tool, 'summarize', os.path.join(scratch, 'prof.ncu-rep'), '--out', args.profile]).returncode
    sys.exit(rc)


class StdoutToStderr(object):
    """the contract is ONE JSON line on stdout: anything a library prints meanwhile (NCCL's version banner under
    NCCL_DEBUG=VERSION, warnings of a C extension) is routed to stderr at the file-descriptor level; `emit` writes the line"""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, text):
        sys.stdout.flush()
        os.write(self.saved, (text + '\n').encode())

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


OUT = None


def emit(text):
    if OUT is not None:
        OUT.emit(text)
    else:
        print(text, flush=True)


def main():
    global OUT
    args = parse()
    if args.profile:
        run_profile(args)
    with StdoutToStderr() as OUT:
        if args.impl == 'reference':
            run_reference(args)
        else:
            run_ours(args)
    OUT = None


if __name__ == '__main__':
    main()
