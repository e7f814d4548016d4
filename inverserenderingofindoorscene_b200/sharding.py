"""Batch sharding of the render path over one process per GPU.

The render layer and the bilateral solver are independent per image, so the batch dimension is
split across ranks with no data-path collective.  The ONE exchange step is the loss normalisation:
``pixelNum`` is the batch-GLOBAL mask sum (wrapperBRDFLight.py:179,192), so exact single-process
parity needs an all-reduce(SUM) of four floats ``[num_render, cnt_render, num_reconst, cnt_reconst]``
(one NCCL call over NVLink; ``gloo`` in the CPU tests), after which every rank scales its local
numerators by ``1/global_cnt``.  The returned losses have the same VALUE on every rank (the global
loss) while their autograd graph only spans the local shard -- calling ``backward()`` on each rank
yields exactly the gradient slices of the single-process run (do NOT average across ranks; under DDP,
which averages parameter gradients, multiply the loss by ``world_size``).
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """contiguous [lo, hi) slice of n items for `rank` (first n % world ranks get one extra)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(tensors, rank=None, world=None):
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    out = []
    for t in tensors:
        lo, hi = shard_range(t.shape[0], rank, world)
        out.append(t[lo:hi])
    return out


def allreduce_loss_terms(terms, group=None):
    """terms: tensor[4] of local sums (detached copy is reduced).  Returns the global sums."""
    red = terms.detach().clone()
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(red, op=dist.ReduceOp.SUM, group=group)
    return red


class _GlobalValueLocalGrad(torch.autograd.Function):
    """value = global_num * scale on every rank; gradient flows through the LOCAL numerator only (d/d local = scale)"""

    @staticmethod
    def forward(ctx, local_num, global_num, scale):
        ctx.save_for_backward(scale)
        return global_num * scale

    @staticmethod
    def backward(ctx, g):
        scale, = ctx.saved_tensors
        return g * scale, None, None


def sharded_losses(render_num, render_cnt, reconst_num, reconst_cnt, envWidth=16, envHeight=8, group=None):
    """render_num = sum((rendered-im)^2*seg) over the LOCAL shard (differentiable), render_cnt = local
    sum(seg); same for the reconstruction pair.  Returns (renderErr, reconstErr) normalised exactly as
    wrapperBRDFLight.py:183-207 does on the full batch.  With one process (no initialised process group or world size
    1) this is the plain single-process normalisation, no collective."""
    dev, dt = render_num.device, render_num.dtype
    multi = dist.is_initialized() and dist.get_world_size(group) > 1
    with torch.no_grad():
        parts = [render_num.detach().reshape(1), torch.as_tensor(render_cnt, device=dev, dtype=dt).reshape(1)]
        if reconst_num is not None:
            parts += [reconst_num.detach().reshape(1), torch.as_tensor(reconst_cnt, device=dev, dtype=dt).reshape(1)]
        g = torch.cat(parts)
        if multi:
            if g.numel() < 4:
                g = torch.cat([g, g.new_zeros(4 - g.numel())])   # always the same 4-float message
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
        cnt = torch.clamp(g[1::2], min=1e-5)
        scale_r = 1.0 / (cnt[0] * 3.0)
        scale_e = 1.0 / (cnt[1] * (3.0 * envWidth * envHeight)) if reconst_num is not None else None
    renderErr = _GlobalValueLocalGrad.apply(render_num, g[0], scale_r)
    reconstErr = None
    if reconst_num is not None:
        reconstErr = _GlobalValueLocalGrad.apply(reconst_num, g[2], scale_e)
    return renderErr, reconstErr
