"""Host-side mirror of the reference's ``BilateralLayer.py``.

  * ``BilateralFunction`` (BilateralLayer.py:20-124): same ``forward(ctx, image, pred, confidence,
    grid_params_arr, bs_params_arr) -> (output, confidence)`` and ``backward -> (None, pred_grad,
    conf_grad, None, None)``; the per-image NumPy/SciPy CPU solve (D2H -> grid build -> PCG -> H2D)
    is replaced by the CUDA solver behind ``irb_bg_*`` (include/irb.h): all images of the batch are
    built and solved together on the device, nothing leaves HBM, and the grid built in forward is
    reused in backward (the reference rebuilds it, BilateralLayer.py:100; results are identical).
  * ``BilateralLayer`` (BilateralLayer.py:126-272): the four hyper-parameter presets, the 12.9 k
    parameter confidence CNN (cuDNN convs; submodule names kept so reference checkpoints load),
    input normalisation, and the call into ``BilateralFunction.apply``.

``cg_maxiter``: the reference reads it from ``bs_params_arr[2]`` (= cg_tol = 1e-5,
BilateralLayer.py:44,79), which current SciPy rejects; this build uses the INTENDED preset
``bs_params_arr[3]`` (12/10/10/10).  ``CG_MAXITER_OVERRIDE`` forces another integer.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib

REQUIRES_CONF_GRAD = True   # BilateralLayer.py:17
CG_MAXITER_OVERRIDE = None  # int to force the PCG iteration count


def _chk(t, name, shape_tail=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ValueError('%s must be a CUDA tensor: this build has no CPU path' % name)
    if t.dtype != torch.float32:
        raise ValueError('%s must be float32 (got %s)' % (name, t.dtype))


def _solver_params(grid_params_arr, bs_params_arr):
    gp = [float(x) for x in grid_params_arr.detach().cpu().tolist()]
    bp = [float(x) for x in bs_params_arr.detach().cpu().tolist()]
    maxiter = int(round(bp[3])) if CG_MAXITER_OVERRIDE is None else int(CG_MAXITER_OVERRIDE)
    return gp[0], gp[1], gp[2], bp[0], bp[1], bp[2], maxiter


class BilateralFunction(torch.autograd.Function):
    """Differentiable bilateral solver: out = slice(argmin_y lam*y'A_s y + (y - splat)' W (...)).

    image [B,3,H,W] guide in [0,1]; pred [B,ch,H,W]; confidence [B,1,H,W];
    grid_params_arr = (sigma_luma, sigma_chroma, sigma_spatial); bs_params_arr = (lam, A_diag_min,
    cg_tol, cg_maxiter).
    """

    @staticmethod
    def forward(ctx, image, pred, confidence, grid_params_arr, bs_params_arr):
        _chk(image, 'image')
        _chk(pred, 'pred')
        _chk(confidence, 'confidence')
        B, ch, H, W = pred.shape
        if tuple(image.shape) != (B, 3, H, W) or tuple(confidence.shape) != (B, 1, H, W):
            raise ValueError('image must be [B,3,H,W] and confidence [B,1,H,W] for pred %s (got %s, %s)'
                             % (tuple(pred.shape), tuple(image.shape), tuple(confidence.shape)))
        sl, sc, ss, lam, dmin, tol, maxiter = _solver_params(grid_params_arr, bs_params_arr)
        dev = pred.device
        image_c, pred_c, conf_c = image.contiguous(), pred.contiguous(), confidence.contiguous()
        lib = _lib.lib()
        nbytes = lib.irb_bg_workspace_bytes(B, H, W, ch)
        if nbytes == 0:
            raise ValueError('unsupported bilateral problem B=%d ch=%d H=%d W=%d' % (B, ch, H, W))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        out = torch.empty(B, ch, H, W, dtype=torch.float32, device=dev)
        yhat = torch.empty(B, ch, H * W, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            _lib.check(lib.irb_bg_build(_lib.ptr(image_c), 255.0, B, H, W, sl, sc, ss, _lib.ptr(ws), nbytes, None, st),
                       'irb_bg_build')
            _lib.check(lib.irb_bg_solve_fwd(_lib.ptr(ws), _lib.ptr(pred_c), _lib.ptr(conf_c), B, ch, H, W, lam, dmin, tol,
                                            maxiter, _lib.ptr(out), _lib.ptr(yhat), st), 'irb_bg_solve_fwd')
        ctx.save_for_backward(pred_c, conf_c)
        ctx.intermediate_results = (ws, yhat)
        ctx.params = (lam, dmin, tol, maxiter)
        return out, confidence

    @staticmethod
    def backward(ctx, grad_output, grad_not_used):
        pred_c, conf_c = ctx.saved_tensors
        ws, yhat = ctx.intermediate_results
        lam, dmin, tol, maxiter = ctx.params
        B, ch, H, W = pred_c.shape
        dev = pred_c.device
        g = grad_output.contiguous()
        pred_grad = torch.empty_like(pred_c)
        conf_grad = torch.empty_like(conf_c) if REQUIRES_CONF_GRAD else None
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().irb_bg_solve_bwd(
                _lib.ptr(ws), _lib.ptr(pred_c), _lib.ptr(conf_c), _lib.ptr(g), _lib.ptr(yhat), B, ch, H, W, lam, dmin,
                tol, maxiter, _lib.ptr(pred_grad), _lib.ptr(conf_grad), _lib.stream_ptr(dev)), 'irb_bg_solve_bwd')
        return None, pred_grad, conf_grad, None, None


# (sigma_luma, sigma_chroma, sigma_spatial), (lam, A_diag_min, cg_tol, cg_maxiter) -- BilateralLayer.py:131-189
_PRESETS = {
    0: ((8, 2, 7), (200, 1e-5, 1e-5, 12)),      # albedo
    1: ((0.5, 0.5, 0.5), (5, 1e-5, 1e-5, 10)),  # normal
    2: ((8, 2, 8), (300, 1e-5, 1e-5, 10)),      # roughness
    4: ((4, 2, 4), (100, 1e-5, 1e-5, 10)),      # depth
}


class BilateralLayer(nn.Module):
    def __init__(self, mode=0, isCuda=True, gpuId=0):
        super(BilateralLayer, self).__init__()
        gp, bp = _PRESETS[mode]
        self.grid_params = {'sigma_luma': gp[0], 'sigma_chroma': gp[1], 'sigma_spatial': gp[2]}
        self.bs_params = {'lam': bp[0], 'A_diag_min': bp[1], 'cg_tol': bp[2], 'cg_maxiter': bp[3]}
        self.grid_params_arr = torch.tensor([float(x) for x in gp], dtype=torch.float32)
        self.bs_params_arr = torch.tensor([float(x) for x in bp], dtype=torch.float32)
        if isCuda:
            self.grid_params_arr = self.grid_params_arr.cuda(gpuId)
            self.bs_params_arr = self.bs_params_arr.cuda(gpuId)
        self.grid_params_arr.requires_grad = False
        self.bs_params_arr.requires_grad = False

        # confidence network: 2 strided convs down, 2 convs up with a skip (BilateralLayer.py:211-232)
        in_ch = 4 if mode in (2, 4) else 6
        self.pad1 = nn.ReplicationPad2d(1)
        self.conv1 = nn.Conv2d(in_channels=in_ch, out_channels=16, kernel_size=4, stride=2, bias=True)
        self.gn1 = nn.GroupNorm(num_groups=2, num_channels=16)
        self.pad2 = nn.ReplicationPad2d(1)
        self.conv2 = nn.Conv2d(in_channels=16, out_channels=16, kernel_size=4, stride=2, bias=True)
        self.gn2 = nn.GroupNorm(num_groups=2, num_channels=16)
        self.dconv1 = nn.Conv2d(in_channels=16, out_channels=16, kernel_size=3, stride=1, padding=1, bias=True)
        self.dgn1 = nn.GroupNorm(num_groups=2, num_channels=16)
        self.dconv2 = nn.Conv2d(in_channels=32, out_channels=16, kernel_size=3, stride=1, padding=1, bias=True)
        self.dgn2 = nn.GroupNorm(num_groups=2, num_channels=16)
        self.dpad3 = nn.ReplicationPad2d(1)
        self.dconvFinal = nn.Conv2d(in_channels=16, out_channels=1, kernel_size=3, stride=1, bias=True)

    def computePadding(self, os, ns):
        assert os <= ns
        gap = ns - os
        return [int((gap + 1) / 2), int((gap - 1) / 2)] if gap % 2 else [int(gap / 2), int(gap / 2)]

    @staticmethod
    def _unit_scale(x):
        # per-sample max over (C,H,W), clamped to [1e-5, 1]  (BilateralLayer.py:244-254)
        s = torch.clamp(x.amax(dim=(1, 2, 3), keepdim=True), 1e-5, 1)
        return x / s

    def forward(self, image, feature, pred):
        image = self._unit_scale(image)
        feature = self._unit_scale(feature)
        inputBatch = torch.cat([image, pred], dim=1).detach()

        x1 = F.relu(self.gn1(self.conv1(self.pad1(inputBatch))), True)
        x2 = F.relu(self.gn2(self.conv2(self.pad2(x1))), True)
        dx1 = F.relu(self.dgn1(self.dconv1(x2)), True)
        dx1 = F.interpolate(dx1, [x1.size(2), x1.size(3)], mode='bilinear')
        dx2 = F.relu(self.dgn2(self.dconv2(torch.cat([dx1, x1], dim=1))), True)
        dx2 = F.interpolate(dx2, [inputBatch.size(2), inputBatch.size(3)], mode='bilinear')
        conf = 0.5 * (torch.tanh(self.dconvFinal(self.dpad3(dx2))) + 1)
        conf = conf / torch.clamp(conf.max(), min=1e-5)

        return BilateralFunction.apply(feature.detach(), pred, conf, self.grid_params_arr, self.bs_params_arr)
