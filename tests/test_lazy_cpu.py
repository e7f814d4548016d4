"""CPU tests of the lazily materialised tensor that carries the env-map between output2env.output2env and
renderingLayer.forwardEnv (inverserenderingofindoorscene_b200/_lazy.py): metadata without materialising, one
materialisation on first use, autograd through the thunk, and the 'untouched' safety net (in-place edits, views, hooks)."""
import io

import torch

from inverserenderingofindoorscene_b200._lazy import LazyTensor


def _make(req=True):
    a = torch.rand(3, 4, requires_grad=req)
    calls = [0]

    def thunk():
        calls[0] += 1
        return a * 2
    return a, calls, LazyTensor(thunk, (3, 4), torch.float32, torch.device('cpu'), req)


def test_metadata_does_not_materialise():
    a, calls, lz = _make()
    assert isinstance(lz, torch.Tensor)
    assert lz.shape == (3, 4) and lz.size() == (3, 4) and lz.size(1) == 4 and lz.dim() == 2 and lz.ndim == 2
    assert lz.dtype == torch.float32 and lz.device.type == 'cpu' and not lz.is_cuda and lz.requires_grad
    assert lz.numel() == 12 and len(lz) == 3 and lz.is_contiguous() and lz.stride() == (4, 1)
    assert 'not materialised' in repr(lz)
    assert calls[0] == 0 and not lz.irb_materialized and lz.irb_untouched()


def test_materialises_once_and_autograd_flows():
    a, calls, lz = _make()
    y = (lz.detach() * 3).sum()
    z = (lz * 3).sum() + torch.log(lz + 1).sum() + lz.view(-1)[0] + torch.cat([lz, lz]).sum() * 0
    assert calls[0] == 1 and lz.irb_materialized and lz.irb_untouched()
    z.backward()
    ref = a.detach().clone().requires_grad_(True)
    r = ref * 2
    ((r * 3).sum() + torch.log(r + 1).sum() + r.view(-1)[0]).backward()
    assert torch.allclose(a.grad, ref.grad)
    assert float(y) == float((a.detach() * 2 * 3).sum())
    # ordinary conversions see the real values
    assert torch.equal(lz.cpu(), (a * 2).detach()) or torch.allclose(lz.cpu(), a * 2)
    assert lz.data.shape == (3, 4) and lz[0].shape == (4,) and lz.expand_as(lz).shape == (3, 4)
    buf = io.BytesIO()
    torch.save(lz.materialize(), buf)


def test_inplace_and_view_writes_are_seen():
    a, calls, lz = _make(req=False)
    lz.mul_(2.0)
    assert not lz.irb_untouched() and torch.allclose(lz.materialize(), a * 4)
    a, calls, lz = _make(req=False)
    v = lz.view(-1)
    assert lz.irb_untouched()
    v[0] = 5.0
    assert not lz.irb_untouched()
    a, calls, lz = _make(req=False)
    lz[1] = 0.0
    assert not lz.irb_untouched()
    a, calls, lz = _make(req=False)
    out = lz
    out *= 2           # rebinding to the real tensor
    assert not isinstance(out, LazyTensor) and not lz.irb_untouched()


def test_gradient_observers_disable_the_fused_dispatch():
    for observe in (lambda t: t.register_hook(lambda g: g), lambda t: t.retain_grad(), lambda t: t.grad):
        a, calls, lz = _make()
        assert lz.irb_untouched()
        observe(lz)
        assert not lz.irb_untouched()


def test_thunk_runs_under_the_producers_grad_mode():
    a = torch.rand(2, 2, requires_grad=True)
    with torch.no_grad():
        lz = LazyTensor(lambda: a * 2, (2, 2), torch.float32, torch.device('cpu'), False)
    assert not lz.materialize().requires_grad
    lz = LazyTensor(lambda: a * 2, (2, 2), torch.float32, torch.device('cpu'), True)
    with torch.no_grad():
        r = lz.materialize()
    assert r.requires_grad
