#!/usr/bin/env python
"""Where does a bench step spend its HOST time?  CPU enqueue time per segment (perf_counter, no syncs) next to the GPU
time per segment (CUDA events), config 2 shapes.  Diagnostic only."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from inverserenderingofindoorscene_b200 import models as M, sharding, _lib  # noqa: E402

B = int(os.environ.get('DIAG_B', '16'))
dev = torch.device('cuda', 0)
host = bench.synth(B, bench.H, bench.W, bench.K, seed=3)
ins = [t.to(dev) for t in host[:6]]
im = torch.rand(B, 3, bench.H, bench.W, device=dev)
seg = (torch.rand(B, 1, bench.H, bench.W, device=dev) > 0.2).float()
layer = M.renderingLayer(imWidth=bench.W, imHeight=bench.H, isCuda=True)
o2e = M.output2env(bench.K, isCuda=True)
names = ['leaves', 'output2env', 'forwardEnv', 'renderLossTerms', 'sharded_losses', 'backward']
N = 30


def step(cpu, evs):
    t = [time.perf_counter()]
    e = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    e[0].record()
    leaves = [x.detach().requires_grad_(True) for x in ins]
    t.append(time.perf_counter()); e[1].record()
    env, _, _, _ = o2e.output2env(leaves[3], leaves[4], leaves[5])
    t.append(time.perf_counter()); e[2].record()
    d, s = layer.forwardEnv(leaves[0], leaves[1], leaves[2], env)
    t.append(time.perf_counter()); e[3].record()
    rend, num, cnt, _ = M.renderLossTerms(d, s, im, seg)
    t.append(time.perf_counter()); e[4].record()
    rerr, _ = sharding.sharded_losses(num, cnt, None, None)
    t.append(time.perf_counter()); e[5].record()
    rerr.backward()
    t.append(time.perf_counter()); e[6].record()
    if cpu is not None:
        for i in range(len(names)):
            cpu[i] += t[i + 1] - t[i]
        evs.append(e)


for _ in range(5):
    step(None, None)
torch.cuda.synchronize()
cpu = [0.0] * len(names)
evs = []
t0 = time.perf_counter()
for _ in range(N):
    step(cpu, evs)
t_enq = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print('B=%d  enqueue %.3f ms/step   wall %.3f ms/step' % (B, t_enq / N * 1e3, t_all / N * 1e3))
for i, n in enumerate(names):
    g = sum(e[i].elapsed_time(e[i + 1]) for e in evs) / N
    print('%-18s cpu %.3f ms   gpu-span %.3f ms' % (n, cpu[i] / N * 1e3, g))
# raw C-ABI call cost
lib = _lib.lib()
t0 = time.perf_counter()
for _ in range(1000):
    lib.irb_launch_count()
print('ctypes call %.2f us' % ((time.perf_counter() - t0) / 1000 * 1e6))
