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
ENV_H, ENV_W = 8, 16


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
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
# CPU arm: the oracle port (restatement of the reference's PyTorch op graph) on host cores
# -------------------------------------------------------------------------------------------

def cpu_step_fn(rows):
    from oracle import render_oracle as ro
    ins = synth(1, rows, W, K, seed=1234 + 1)
    leaves = [t.clone().requires_grad_(True) for t in ins[:6]]
    gD, gS = ins[6], ins[7]
    # the full-frame camera rays, cropped to the sampled rows (pixels are independent)
    v_full = torch.from_numpy(ro.camera_rays(H, W))[:, :rows].contiguous()

    def render(alb, nrm, rg, ax, lam, wt):
        env, _, _, _ = ro.output2env(ax, lam, wt)
        return ro.render_env(alb, nrm, rg, env, view=v_full)

    def step():
        d, s = render(*leaves)
        torch.autograd.grad((d * gD).sum() + (s * gS).sum(), leaves)
    return step


def _best_threads(cores):
    """torch CPU elementwise ops stop scaling (and can collapse) at high thread counts on medium tensors:
    probe a 48-row sample at a few thread counts and keep the fastest -- the baseline gets its best setting."""
    best, best_t = None, 1e30
    step = cpu_step_fn(48)
    for n in sorted(set([min(cores, x) for x in (4, 8, 16, 32, 64, cores)])):
        torch.set_num_threads(n)
        step()
        t0 = time.perf_counter()
        step()
        t = time.perf_counter() - t0
        if t < best_t:
            best, best_t = n, t
        elif t > 1.25 * best_t:
            break  # past the scaling knee: more threads only get slower (they collapse at 100+ threads)
    torch.set_num_threads(best)
    return best, best_t / 48


def cpu_baseline(budget_s=20.0, rows=0):
    """Time the oracle port on a bounded sample (rows x 320 pixels of one 240x320 frame)."""
    cores = os.cpu_count() or 1
    threads, per_row = _best_threads(cores)
    if rows <= 0:
        rows = int(max(48, min(H, budget_s / 3.0 / max(per_row, 1e-6))))
    step = cpu_step_fn(rows)
    step()  # warm-up
    ts = []
    for _ in range(2):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    t = sorted(ts)[0]
    return {'value': rows * W / t / 1e6, 'unit': 'Mpix/s', 'cores': threads, 'kind': 'port',
            'sample': 'oracle/render_oracle.py fp32 fwd+bwd, %d x %d pixels of one 240x320 frame, 12 SG, '
                      'torch %s CPU, %d threads (best of a probe over <=%d host cores), best of 2 after 1 warm-up'
                      % (rows, W, torch.__version__, threads, cores)}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads, per_row = _best_threads(cores)
    # size the per-step sample so that the (steps+warmup) steps end within ~90 s: first guess from the thread probe,
    # then one calibration step at that size (the probe's small samples under-estimate the per-row cost by ~2x)
    nstep = max(1, args.steps + args.warmup)
    rows = int(max(8, min(H, 90.0 / nstep / max(per_row, 1e-6))))
    step = cpu_step_fn(rows)
    t0 = time.perf_counter()
    step()
    per_row = (time.perf_counter() - t0) / rows
    rows2 = int(max(8, min(H, 90.0 / nstep / max(per_row, 1e-6))))
    if rows2 != rows:
        rows = rows2
        step = cpu_step_fn(rows)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = rows * W * args.steps / dt / 1e6
    sample = ('oracle port (restatement of models.output2env + renderingLayer.forwardEnv + autograd), fp32, '
              '%d x %d pixels of one 240x320 frame per step, %d threads (best of a probe over <=%d host cores)'
              % (rows, W, threads, cores))
    print(json.dumps({
        'impl': 'reference', 'metric': 'Mpixels/s render-layer fwd+bwd (240x320, 12 SG)', 'value': val,
        'unit': 'Mpix/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'config2: fused SG render fwd+bwd, 240x320, 12 SG (CPU sample: %d rows)' % rows},
        'cpu_baseline': {'value': val, 'unit': 'Mpix/s', 'cores': threads, 'kind': 'port', 'sample': sample},
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

def run_ours(args):
    import torch.distributed as dist
    from inverserenderingofindoorscene_b200 import models as irb_models
    from inverserenderingofindoorscene_b200 import _lib

    _lib.lib()  # fail loudly if the CUDA library is missing
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device (there is no CPU path for --impl ours)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    B = B_PER_GPU
    host = synth(B, H, W, K, seed=1234 + 2 + rank, pinned=True)
    devt = [t.to(dev) for t in host]
    layer = irb_models.renderingLayer(imWidth=W, imHeight=H, isCuda=True)
    gD, gS = devt[6], devt[7]
    npx = B * H * W

    def step_device(ev=None):
        leaves = [t.detach().requires_grad_(True) for t in devt[:6]]
        if ev:
            ev[0].record()
        d, s = layer.forwardSG(*leaves)
        if ev:
            ev[1].record()
        torch.autograd.backward([d, s], [gD, gS])
        if ev:
            ev[2].record()
        return d, s, leaves

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- value: device-resident inputs -------------------------------------------------------
    for _ in range(max(3, args.warmup)):
        step_device()
    sync_all()
    clocks = Clocks(local)
    clocks.start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for i in range(args.steps):
        step_device(evs[i])
    e1.record()
    sync_all()
    clocks.stop_flag = True
    ms_total = e0.elapsed_time(e1)
    fwd_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    bwd_ms = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * npx / (ms_step * 1e-3) / 1e6

    # ---- e2e: pinned host buffers -> public API -> loss + rendered image back to host ---------
    # Every step copies ITS inputs host->device and reads ITS results back; like any production feeder the copies
    # of step i+1 run on a second stream while step i computes (double-buffered device inputs).
    im = torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(99)).pin_memory()
    out_host = torch.empty(B, 3, H, W).pin_memory()
    loss_host = torch.empty(4).pin_memory()
    npx_t = torch.tensor(float(npx), device=dev)
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [[torch.empty_like(h, device=dev) for h in host[:6]] + [torch.empty_like(im, device=dev)] for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    freed = [torch.cuda.Event(), torch.cuda.Event()]

    def upload(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[slot])          # the compute stream is done with this slot
            for dst, src in zip(bufs[slot], list(host[:6]) + [im]):
                dst.copy_(src, non_blocking=True)
            ready[slot].record(copy_stream)

    def step_e2e(slot, prefetch):
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(ready[slot])
        if prefetch:
            upload(slot ^ 1)
        ins = [t.detach().requires_grad_(True) for t in bufs[slot][:6]]
        imd = bufs[slot][6]
        d, s = layer.forwardSG(*ins)
        rend = torch.clamp(d + s, 0, 1)
        diff = rend - imd
        part = torch.stack([(diff * diff).sum(), npx_t, d.detach().sum(), s.detach().sum()])
        if world > 1:
            red = part.detach().clone()
            dist.all_reduce(red)           # the path's one exchange step: 4-float loss all-reduce
            loss = part[0] / red[1] / 3.0
        else:
            red = part.detach()
            loss = part[0] / part[1] / 3.0
        loss.backward()
        freed[slot].record(cur)
        out_host.copy_(rend.detach(), non_blocking=True)
        loss_host.copy_(red, non_blocking=True)
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
    e2e_ms = e0.elapsed_time(e1) / n_e2e
    t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    h2d = sum(h.numel() * 4 for h in host[:6]) + im.numel() * 4
    d2h = out_host.numel() * 4 + loss_host.numel() * 4

    if rank == 0:
        peak, peak_src = peaks()
        bwd_bytes = (20 + 14 * K) * 4 * npx      # reads 13+7K, writes 7K+7 floats per pixel
        fwd_bytes = (13 + 7 * K) * 4 * npx       # reads 7+7K, writes 6
        ach = bwd_bytes / (bwd_ms * 1e-3) / 1e9
        clk = clocks.summary()
        sm_mhz = clk.get('sm_mhz') or 1965
        # instruction model read off the SASS of the two kernels (profiles/r01): FMA-pipe instructions per pixel
        # and direction -- forward 12 lobes x 5 + ~50 BRDF = 110, backward 12 x 13 + ~95 = 250; 2 FLOP each
        flop_fwd = 128 * 110 * 2.0 * npx
        flop_bwd = 128 * 250 * 2.0 * npx
        fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6
        # DRAM traffic of the backward kernel from the committed ncu --set full capture of this config
        # (profiles/r01/ncu_v9_bwd_summary.txt: dram read 483.7 MB + write 407.2 MB per launch)
        traffic_bwd = 483.660800e6 + 407.164416e6 if (B, H, W, K) == (16, 240, 320, 12) else None
        line = {
            'metric': 'Mpixels/s render-layer fwd+bwd (240x320, 12 SG)', 'value': value, 'unit': 'Mpix/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup), 'ms_per_step': ms_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'config2: fused SG render fwd+bwd (models.renderingLayer.forwardSG), '
                                   'batch 16 per GPU x 240x320, 12 SG, all input grads',
                       'global_batch': B * world, 'grid': [H, W], 'sg_lobes': K, 'parallelism': 'batch-shard x%d' % world,
                       'l2': 'inputs per step (447 MB/GPU) exceed the 126 MB L2; no explicit flush'},
            'kernel_ms': {'render_sg_fwd': fwd_ms, 'render_sg_bwd': bwd_ms},
            'roofline': {'bound': 'hbm', 'kernel': 'render_sg_group_kernel<1,2,1,1,2> (fused backward, lobe-split)', 'achieved': ach,
                         'peak': peak, 'unit': 'GB/s', 'frac': ach / peak, 'traffic': traffic_bwd,
                         'traffic_source': 'profiles/r01/ncu_v9_bwd_summary.txt: ncu --set full, dram__bytes_read+write per launch',
                         'peak_source': peak_src,
                         'algorithmic_bytes_per_launch': bwd_bytes,
                         'fwd': {'achieved': fwd_bytes / (fwd_ms * 1e-3) / 1e9, 'algorithmic_bytes_per_launch': fwd_bytes},
                         'fp32': {'note': 'kernel is FP32-FMA/MUFU bound, not HBM bound (DESIGN.md 4.1): FMA-pipe '
                                          'instruction model x 2 FLOP vs 148 SM x 128 lanes x 2 x sm clock',
                                  'bwd_tflops': flop_bwd / (bwd_ms * 1e-3) / 1e12,
                                  'fwd_tflops': flop_fwd / (fwd_ms * 1e-3) / 1e12,
                                  'peak_tflops': fp32_peak / 1e12,
                                  'bwd_frac': flop_bwd / (bwd_ms * 1e-3) / fp32_peak,
                                  'fwd_frac': flop_fwd / (fwd_ms * 1e-3) / fp32_peak}},
            'e2e': {'value': world * npx / (e2e_ms * 1e-3) / 1e6, 'unit': 'Mpix/s', 'ms_per_step': e2e_ms,
                    'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'steps': n_e2e},
            'gpu_launches': 2 * args.steps,   # one lobe-split kernel for the forward, one for the backward
            'clocks': clk,
        }
        if not args.no_cpu_baseline and world == 1:
            line['cpu_baseline'] = cpu_baseline(rows=args.cpu_baseline_rows)
        elif world > 1:
            line['cpu_baseline'] = {'value': None, 'unit': 'Mpix/s', 'cores': os.cpu_count(), 'kind': 'port',
                                    'sample': 'reported at N=1 only'}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
