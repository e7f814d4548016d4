"""a19: BilateralLayer shell parity against the reference's own module (baseline/_ref, unmodified) with ONE shared
state_dict: the normalised guide, the confidence map and the parameter arrays that reach BilateralFunction.apply
(BilateralLayer.py:244-272) must agree.  Runs on the CPU (the shell is PyTorch: normalisation + conf-net); the solver call
is intercepted on both sides, so neither SciPy nor the CUDA library is involved.  A GPU variant runs under -m gpu."""
import pytest
import torch

from conftest import relerr


def _ref_or_skip(*names):
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip('baseline/_ref not present (run __graft_entry__.build() in the build container)')
    return ref_loader.load(*names)


def _shell(mode, dev):
    from inverserenderingofindoorscene_b200 import BilateralLayer as BL
    ref_bl = _ref_or_skip('BilateralLayer')
    torch.manual_seed(mode)
    cuda = dev.type == 'cuda'
    ours = BL.BilateralLayer(mode=mode, isCuda=cuda).to(dev)
    ref = ref_bl.BilateralLayer(mode=mode, isCuda=cuda, gpuId=0).to(dev)
    missing = ref.load_state_dict(ours.state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    ch = 3 if mode == 0 else 1
    g = torch.Generator(device=dev).manual_seed(5)
    im = torch.rand(2, 3, 64, 96, device=dev, generator=g) * 1.3
    feat = torch.rand(2, 3, 64, 96, device=dev, generator=g) * 0.8
    pred = torch.rand(2, ch, 64, 96, device=dev, generator=g)
    captured = {}

    def grab(key):
        class _Stub(torch.autograd.Function):
            @staticmethod
            def forward(ctx, image, p, conf, gp, bp):
                captured[key] = (image.detach().clone(), conf.detach().clone(), gp.detach().clone(), bp.detach().clone())
                return p.clone(), conf.clone()

            @staticmethod
            def backward(ctx, a, b):
                return None, a, b, None, None
        return _Stub
    saved_ref, saved_ours = ref_bl.BilateralFunction, BL.BilateralFunction
    ref_bl.BilateralFunction, BL.BilateralFunction = grab('ref'), grab('ours')
    try:
        ref(im, feat, pred)
        ours(im, feat, pred)
    finally:
        ref_bl.BilateralFunction, BL.BilateralFunction = saved_ref, saved_ours
    for a, b, name in zip(captured['ours'], captured['ref'], ('guide', 'conf', 'grid_params', 'bs_params')):
        assert a.shape == b.shape, name
        assert relerr(a.cpu().numpy(), b.cpu().numpy()) <= 1e-5, name
    assert float(captured['ours'][1].max()) == pytest.approx(1.0, abs=1e-6)


@pytest.mark.parametrize('mode', [0, 2, 4])
def test_bilateral_layer_shell_vs_reference_cpu(mode):
    _shell(mode, torch.device('cpu'))


@pytest.mark.gpu
@pytest.mark.parametrize('mode', [0, 2, 4])
def test_bilateral_layer_shell_vs_reference_gpu(mode):
    _shell(mode, torch.device('cuda'))
