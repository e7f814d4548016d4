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


def _global_value_local_grad(local_num, global_num, scale):
    # value = global_num*scale on every rank; gradient flows through the local numerator only
    return (local_num - local_num.detach() + global_num) * scale


def sharded_losses(render_num, render_cnt, reconst_num, reconst_cnt, envWidth=16, envHeight=8, group=None):
    """render_num = sum((rendered-im)^2*seg) over the LOCAL shard (differentiable), render_cnt = local
    sum(seg); same for the reconstruction pair.  Returns (renderErr, reconstErr) normalised exactly as
    wrapperBRDFLight.py:183-207 does on the full batch."""
    dev = render_num.device
    z = torch.zeros((), device=dev, dtype=render_num.dtype)
    terms = torch.stack([render_num.detach(), torch.as_tensor(render_cnt, device=dev, dtype=render_num.dtype),
                         reconst_num.detach() if reconst_num is not None else z,
                         torch.as_tensor(reconst_cnt if reconst_cnt is not None else 0.0, device=dev,
                                         dtype=render_num.dtype)])
    g = allreduce_loss_terms(terms, group)
    cnt_r = torch.clamp(g[1], min=1e-5)
    renderErr = _global_value_local_grad(render_num, g[0], 1.0 / cnt_r / 3.0)
    reconstErr = None
    if reconst_num is not None:
        cnt_e = torch.clamp(g[3], min=1e-5)
        reconstErr = _global_value_local_grad(reconst_num, g[2], 1.0 / cnt_e / 3.0 / envWidth / envHeight)
    return renderErr, reconstErr
