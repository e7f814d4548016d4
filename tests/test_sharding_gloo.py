"""world_size-2 gloo test of the N>1 host logic: batch sharding + the 4-float loss all-reduce reproduce
the single-process loss value and gradients exactly.  The render outputs are stood in by
a differentiable per-image function (the CUDA kernels need a GPU; this tests the sharding arithmetic)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_render(x):
    # any per-image differentiable map (images are independent in the real layer too)
    return torch.tanh(x) * x.mean(dim=(1, 2, 3), keepdim=True)


def _losses(x, im, seg, env, envgt, segenv, sharded, sharding):
    r = torch.clamp(_fake_render(x), 0, 1)
    num_r = ((r - im) * (r - im) * seg).sum()
    cnt_r = seg.sum()
    d = torch.log(env + 1.0) - torch.log(envgt + 1.0)
    num_e = (d * d * segenv).sum()
    cnt_e = segenv.sum()
    if sharded:
        return sharding.sharded_losses(num_r, cnt_r, num_e, cnt_e)
    return num_r / max(float(cnt_r), 1e-5) / 3.0, num_e / max(float(cnt_e), 1e-5) / 3.0 / 16 / 8


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from inverserenderingofindoorscene_b200 import sharding
    g = torch.Generator().manual_seed(0)
    B = 5  # not divisible by 2: ragged shards
    x = torch.randn(B, 3, 6, 8, generator=g, dtype=torch.float64)
    im = torch.rand(B, 3, 6, 8, generator=g, dtype=torch.float64)
    seg = (torch.rand(B, 1, 6, 8, generator=g) > 0.3).double()
    e = torch.rand(B, 3, 6, 8, 4, generator=g, dtype=torch.float64)
    egt = torch.rand(B, 3, 6, 8, 4, generator=g, dtype=torch.float64)
    sege = (torch.rand(B, 1, 6, 8, 1, generator=g) > 0.3).double()
    # single-process truth
    xf = x.clone().requires_grad_(True)
    ef = e.clone().requires_grad_(True)
    lr, le = _losses(xf, im, seg, ef, egt, sege, False, sharding)
    (lr + 10 * le).backward()
    # sharded
    xs, ims, segs, es, egts, seges = sharding.shard_batch([x, im, seg, e, egt, sege], rank, world)
    xs = xs.clone().requires_grad_(True)
    es = es.clone().requires_grad_(True)
    slr, sle = _losses(xs, ims, segs, es, egts, seges, True, sharding)
    (slr + 10 * sle).backward()
    lo, hi = sharding.shard_range(B, rank, world)
    ok = (abs(float(slr) - float(lr)) < 1e-12 and abs(float(sle) - float(le)) < 1e-12 and
          torch.allclose(xs.grad, xf.grad[lo:hi], atol=1e-14) and torch.allclose(es.grad, ef.grad[lo:hi], atol=1e-14))
    q.put((rank, bool(ok), hi - lo))
    dist.destroy_process_group()


def test_sharded_loss_matches_single_process():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert [r[1] for r in res] == [True, True]
    assert sum(r[2] for r in res) == 5


def test_shard_range_covers_everything():
    from inverserenderingofindoorscene_b200 import sharding
    for n in (1, 5, 16, 33):
        for w in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
