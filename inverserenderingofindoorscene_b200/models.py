"""Host-side mirror of the hot-path classes of the reference's ``models.py``.

Same names, constructor keywords, public attributes, method names and return values as
``/root/reference/models.py``, so ``wrapperBRDFLight.py``, ``trainLight.py``,
``testReal.py`` ... call them unchanged:

  * ``output2env``      (models.py:348-404)  -> CUDA kernels ``irb_sg_to_env_fwd/bwd``
  * ``renderingLayer``  (models.py:407-522)  -> ``irb_render_env_fwd/bwd`` or, when the env-map
                                               was produced by ``output2env`` (the trainLight /
                                               testReal call pattern), the FUSED
                                               ``irb_render_sg_fwd/bwd`` that never reads it
  * ``LSregress`` / ``LSregressDiffSpec`` (models.py:7-84) -- per-sample scale regressions, PyTorch

Everything numeric in the two classes runs in hand-written sm_100a kernels behind the C ABI
(include/irb.h) through ctypes; PyTorch only owns memory, streams and the autograd graph.
There is no CPU path: CPU tensors raise.
"""
import numpy as np
import torch

from . import _lib
from ._lazy import LazyTensor, unlazy

__all__ = ['output2env', 'renderingLayer', 'LSregress', 'LSregressDiffSpec', 'renderLoss', 'renderLossTerms']


# A/B and test knob: False makes the fused backward rebuild U_k / V_k itself (no side tensor from the forward)
_SAVE_LOBE_SUMS = True
# measurement hook: a list that receives ('render_sg_bwd', start_event, end_event) per fused backward launch (bench.py)
_EVENT_SINK = None


def _chk(t, name, ndim=None):
    if not isinstance(t, torch.Tensor):
        raise ValueError('%s must be a torch.Tensor' % name)
    if not t.is_cuda:
        raise ValueError('%s must be a CUDA tensor: this build has no CPU path (got device %s)' % (name, t.device))
    if t.dtype != torch.float32:
        raise ValueError('%s must be float32 (got %s)' % (name, t.dtype))
    if ndim is not None and t.dim() != ndim:
        raise ValueError('%s must have %d dims (got shape %s)' % (name, ndim, tuple(t.shape)))


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


# ------------------------------------------------------------------------------------------
# autograd bridges to the C ABI
# ------------------------------------------------------------------------------------------

class _AvgPool(torch.autograd.Function):
    """F.adaptive_avg_pool2d (models.py:465-469) via irb_avgpool_fwd/bwd."""

    @staticmethod
    def forward(ctx, x, R, C):
        x = _c(x)
        B, ch, H, W = x.shape
        ctx.dims = (B * ch, H, W, R, C)
        out = torch.empty(B, ch, R, C, device=x.device, dtype=x.dtype)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().irb_avgpool_fwd(_lib.ptr(x), _lib.ptr(out), B * ch, H, W, R, C,
                                                  _lib.stream_ptr(x.device)), 'irb_avgpool_fwd')
        return out

    @staticmethod
    def backward(ctx, g):
        N, H, W, R, C = ctx.dims
        g = _c(g)
        gi = torch.empty(g.shape[0], g.shape[1], H, W, device=g.device, dtype=g.dtype)
        with torch.cuda.device(g.device):
            _lib.check(_lib.lib().irb_avgpool_bwd(_lib.ptr(g), _lib.ptr(gi), N, H, W, R, C,
                                                  _lib.stream_ptr(g.device)), 'irb_avgpool_bwd')
        return gi, None, None


class _AvgPool3(torch.autograd.Function):
    """albedo / normal / rough pooled to the env grid in ONE launch each way (irb_avgpool3_fwd/bwd, models.py:465-469)."""

    @staticmethod
    def forward(ctx, a, n, r, R, C):
        a, n, r = _c(a), _c(n), _c(r)
        B, _, H, W = a.shape
        outs = [torch.empty(B, t.shape[1], R, C, device=a.device, dtype=a.dtype) for t in (a, n, r)]
        with torch.cuda.device(a.device):
            _lib.check(_lib.lib().irb_avgpool3_fwd(_lib.ptr(a), _lib.ptr(n), _lib.ptr(r), *[_lib.ptr(o) for o in outs],
                                                   B * a.shape[1], B * n.shape[1], B * r.shape[1], H, W, R, C,
                                                   _lib.stream_ptr(a.device)), 'irb_avgpool3_fwd')
        ctx.dims = (B, a.shape[1], n.shape[1], r.shape[1], H, W, R, C)
        return tuple(outs)

    @staticmethod
    def backward(ctx, ga, gn, gr):
        B, ca, cn, cr, H, W, R, C = ctx.dims
        need = ctx.needs_input_grad
        gs = [(_c(g) if (g is not None and need[i]) else None) for i, g in enumerate((ga, gn, gr))]
        ref = next((g for g in gs if g is not None), None)
        if ref is None:
            return None, None, None, None, None
        gi = [torch.empty(B, c, H, W, device=ref.device, dtype=ref.dtype) if g is not None else None
              for g, c in zip(gs, (ca, cn, cr))]
        with torch.cuda.device(ref.device):
            _lib.check(_lib.lib().irb_avgpool3_bwd(*[_lib.ptr(g) for g in gs], *[_lib.ptr(g) for g in gi], B * ca, B * cn, B * cr,
                                                   H, W, R, C, _lib.stream_ptr(ref.device)), 'irb_avgpool3_bwd')
        return gi[0], gi[1], gi[2], None, None


def _pool3(a, n, r, R, C):
    """the three BRDF maps at the env-grid resolution (identity when they already are)"""
    if a.shape[2:] == n.shape[2:] == r.shape[2:]:
        if a.shape[2] == R and a.shape[3] == C:
            return a, n, r
        return _AvgPool3.apply(a, n, r, R, C)
    return _pool(a, R, C), _pool(n, R, C), _pool(r, R, C)


def _pool(x, R, C):
    if x.shape[2] == R and x.shape[3] == C:
        return x
    return _AvgPool.apply(x, R, C)


class _SgToEnv(torch.autograd.Function):
    """models.py:371-404.  Returns (env, lamb_w, weight_w)."""

    @staticmethod
    def forward(ctx, axis, lamb, weight, envH, envW, warp):
        axis, lamb, weight = _c(axis), _c(lamb), _c(weight)
        B, K, _, R, C = axis.shape
        dev = axis.device
        env = torch.empty(B, 3, R, C, envH, envW, device=dev, dtype=torch.float32)
        if warp:
            lamb_w = torch.empty_like(lamb)
            weight_w = torch.empty_like(weight)
        else:
            lamb_w = weight_w = None
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().irb_sg_to_env_fwd(
                _lib.ptr(axis), _lib.ptr(lamb), _lib.ptr(weight), _lib.ptr(env), _lib.ptr(lamb_w),
                _lib.ptr(weight_w), B, K, R, C, envH, envW, int(warp), _lib.stream_ptr(dev)), 'irb_sg_to_env_fwd')
        ctx.save_for_backward(axis, lamb, weight, lamb_w, weight_w)
        ctx.cfg = (envH, envW, int(warp))
        ctx.set_materialize_grads(False)
        if warp:
            return env, lamb_w, weight_w
        return env, lamb, weight

    @staticmethod
    def backward(ctx, g_env, g_lw, g_ww):
        axis, lamb, weight, lamb_w, weight_w = ctx.saved_tensors
        envH, envW, warp = ctx.cfg
        B, K, _, R, C = axis.shape
        dev = axis.device
        need = ctx.needs_input_grad
        g_axis = g_lamb = g_weight = None
        if g_env is not None and any(need[:3]):
            g_env = _c(g_env)
            g_axis = torch.empty_like(axis) if need[0] else None
            g_lamb = torch.empty_like(lamb) if need[1] else None
            g_weight = torch.empty_like(weight) if need[2] else None
            with torch.cuda.device(dev):
                _lib.check(_lib.lib().irb_sg_to_env_bwd(
                    _lib.ptr(axis), _lib.ptr(lamb), _lib.ptr(weight), _lib.ptr(g_env), _lib.ptr(g_axis),
                    _lib.ptr(g_lamb), _lib.ptr(g_weight), B, K, R, C, envH, envW, warp, _lib.stream_ptr(dev)),
                    'irb_sg_to_env_bwd')
        # gradients arriving through the (rarely used) warped lamb / weight outputs
        cw = np.pi / 2 * 0.999
        if g_lw is not None and need[1]:
            t = g_lw * (cw * (1 + lamb_w * lamb_w)) if warp else g_lw
            g_lamb = t if g_lamb is None else g_lamb + t
        if g_ww is not None and need[2]:
            t = g_ww * (cw * (1 + weight_w * weight_w)) if warp else g_ww
            g_weight = t if g_weight is None else g_weight + t
        return g_axis, g_lamb, g_weight, None, None, None


class _SgReconstLoss(torch.autograd.Function):
    """sum((log(c*E(sg) + off) - log(envGT + off))^2 * seg) with c the per-sample LSregress scale, the env-map E never
    leaving the kernels (wrapperBRDFLight.py:172-188).  Returns (numerator, sum(seg), coef[B])."""

    @staticmethod
    def forward(ctx, axis, lamb, weight, env_gt, seg, envH, envW, offset):
        axis, lamb, weight, env_gt, seg = _c(axis), _c(lamb), _c(weight), _c(env_gt), _c(seg)
        B, K, _, R, C = axis.shape
        dev = axis.device
        stats = torch.empty(B, 4, device=dev, dtype=torch.float64)
        # the SG gradients are small (7K floats per pixel against the 384 of the env-map): when they are wanted the
        # second pass computes them together with the loss for a unit upstream gradient, and backward only scales them
        need = ctx.needs_input_grad
        unit = [torch.empty_like(t) if need[i] else None for i, t in enumerate((axis, lamb, weight))]
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().irb_sg_reconst_fwd(
                _lib.ptr(axis), _lib.ptr(lamb), _lib.ptr(weight), _lib.ptr(env_gt), _lib.ptr(seg), _lib.ptr(stats),
                *[_lib.ptr(u) for u in unit], B, K, R, C, envH, envW, 1, float(offset), _lib.stream_ptr(dev)),
                'irb_sg_reconst_fwd')
        ctx.unit = unit
        tot = stats.sum(dim=0)
        coef = torch.clamp(stats[:, 0] / torch.clamp(stats[:, 1], min=1e-5), 0.001, 1000).float()
        num, cnt = tot[2].float(), tot[3].float()
        ctx.mark_non_differentiable(cnt, coef)
        return num, cnt, coef

    @staticmethod
    def backward(ctx, g_num, g_cnt, g_coef):
        g = g_num.detach().to(torch.float32)
        return tuple(None if u is None else u * g for u in ctx.unit) + (None,) * 5


class _RenderEnv(torch.autograd.Function):
    """models.py:466-520 on pooled maps and a materialised env-map."""

    @staticmethod
    def forward(ctx, albedo, normal, rough, env, view, F0):
        albedo, normal, rough, env = _c(albedo), _c(normal), _c(rough), _c(env)
        B, _, R, C, envH, envW = env.shape
        dev = env.device
        od = torch.empty(B, 3, R, C, device=dev, dtype=torch.float32)
        os_ = torch.empty_like(od)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().irb_render_env_fwd(
                _lib.ptr(albedo), _lib.ptr(normal), _lib.ptr(rough), _lib.ptr(env), _lib.ptr(view), _lib.ptr(od),
                _lib.ptr(os_), B, R, C, envH, envW, float(F0), _lib.stream_ptr(dev)), 'irb_render_env_fwd')
        ctx.save_for_backward(albedo, normal, rough, env, view)
        ctx.F0 = float(F0)
        return od, os_

    @staticmethod
    def backward(ctx, g_d, g_s):
        albedo, normal, rough, env, view = ctx.saved_tensors
        B, _, R, C, envH, envW = env.shape
        dev = env.device
        need = ctx.needs_input_grad
        g_d = _c(g_d) if g_d is not None else torch.zeros(B, 3, R, C, device=dev)
        g_s = _c(g_s) if g_s is not None else torch.zeros(B, 3, R, C, device=dev)
        ga = torch.empty_like(albedo) if need[0] else None
        gn = torch.empty_like(normal) if need[1] else None
        gr = torch.empty_like(rough) if need[2] else None
        ge = torch.empty_like(env) if need[3] else None
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().irb_render_env_bwd(
                _lib.ptr(albedo), _lib.ptr(normal), _lib.ptr(rough), _lib.ptr(env), _lib.ptr(view), _lib.ptr(g_d),
                _lib.ptr(g_s), _lib.ptr(ga), _lib.ptr(gn), _lib.ptr(gr), _lib.ptr(ge), B, R, C, envH, envW, ctx.F0,
                _lib.stream_ptr(dev)), 'irb_render_env_bwd')
        return ga, gn, gr, ge, None, None


class _RenderSG(torch.autograd.Function):
    """Fused output2env + forwardEnv (wrapperBRDFLight.py:177,194) on pooled maps."""

    @staticmethod
    def forward(ctx, albedo, normal, rough, axis, lamb, weight, view, F0, envH, envW, warp):
        albedo, normal, rough = _c(albedo), _c(normal), _c(rough)
        axis, lamb, weight = _c(axis), _c(lamb), _c(weight)
        B, K, _, R, C = axis.shape
        dev = axis.device
        od = torch.empty(B, 3, R, C, device=dev, dtype=torch.float32)
        os_ = torch.empty_like(od)
        # the per-lobe hemisphere sums U_k, V_k: written by the forward when a lobe-side gradient will be asked for, so that
        # the backward does not rebuild them (96 B/pixel at K = 12 on a path that uses 7 % of the HBM bandwidth)
        need = ctx.needs_input_grad
        uv = (torch.empty(B, 2 * K, R, C, device=dev, dtype=torch.float32)
              if (_SAVE_LOBE_SUMS and (need[0] or any(need[3:6]))) else None)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().irb_render_sg_fwd(
                _lib.ptr(albedo), _lib.ptr(normal), _lib.ptr(rough), _lib.ptr(axis), _lib.ptr(lamb),
                _lib.ptr(weight), _lib.ptr(view), _lib.ptr(od), _lib.ptr(os_), _lib.ptr(uv), B, K, R, C, envH, envW,
                float(F0), int(warp), _lib.stream_ptr(dev)), 'irb_render_sg_fwd')
        ctx.save_for_backward(albedo, normal, rough, axis, lamb, weight, view)
        ctx.uv = uv
        ctx.cfg = (float(F0), envH, envW, int(warp))
        return od, os_

    @staticmethod
    def backward(ctx, g_d, g_s):
        albedo, normal, rough, axis, lamb, weight, view = ctx.saved_tensors
        F0, envH, envW, warp = ctx.cfg
        B, K, _, R, C = axis.shape
        dev = axis.device
        need = ctx.needs_input_grad
        g_d = _c(g_d) if g_d is not None else torch.zeros(B, 3, R, C, device=dev)
        g_s = _c(g_s) if g_s is not None else torch.zeros(B, 3, R, C, device=dev)
        outs = [torch.empty_like(t) if need[i] else None
                for i, t in enumerate((albedo, normal, rough, axis, lamb, weight))]
        sink = _EVENT_SINK
        with torch.cuda.device(dev):
            if sink is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            _lib.check(_lib.lib().irb_render_sg_bwd(
                _lib.ptr(albedo), _lib.ptr(normal), _lib.ptr(rough), _lib.ptr(axis), _lib.ptr(lamb),
                _lib.ptr(weight), _lib.ptr(view), _lib.ptr(g_d), _lib.ptr(g_s), _lib.ptr(ctx.uv),
                *[_lib.ptr(o) for o in outs], B, K, R, C, envH, envW, F0, warp, _lib.stream_ptr(dev)), 'irb_render_sg_bwd')
            if sink is not None:
                e1.record()
                sink.append(('render_sg_bwd', e0, e1))
        return tuple(outs) + (None, None, None, None, None)


# ------------------------------------------------------------------------------------------
# public classes (reference names)
# ------------------------------------------------------------------------------------------

def _direction_table(envWidth, envHeight):
    """ls[d] = (sinEl cosAz, sinEl sinAz, cosEl), d = i*envWidth + j (models.py:353-362, 437-446)."""
    az = ((np.arange(envWidth) + 0.5) / envWidth - 0.5) * 2 * np.pi
    el = ((np.arange(envHeight) + 0.5) / envHeight) * np.pi / 2.0
    az, el = np.meshgrid(az, el)
    return np.sin(el) * np.cos(az), np.sin(el) * np.sin(az), np.cos(el), el


class output2env():
    """SG lobes -> env-map.  Mirror of models.output2env (models.py:348-404)."""

    MAX_FUSED_LOBES = 24

    def __init__(self, SGNum, envWidth=16, envHeight=8, isCuda=True):
        self.envWidth = envWidth
        self.envHeight = envHeight
        lx, ly, lz, _ = _direction_table(envWidth, envHeight)
        ls = np.stack([lx, ly, lz], axis=0)[np.newaxis, np.newaxis, :, np.newaxis, np.newaxis, :, :]
        self.ls = torch.from_numpy(ls.astype(np.float32))
        self.SGNum = SGNum
        self.isCuda = isCuda
        if isCuda:
            self.ls = self.ls.cuda()
        self.ls.requires_grad = False

    def _check(self, axis, lamb, weight):
        if any(isinstance(t, LazyTensor) for t in (axis, lamb, weight)):
            raise ValueError('axis / lamb / weight must be ordinary tensors (got a LazyTensor; pass unlazy(t))')
        _chk(axis, 'axis', 5)
        _chk(lamb, 'lamb', 4)
        _chk(weight, 'weight', 4)
        B, K, three, R, C = axis.shape
        if K != self.SGNum or three != 3:
            raise ValueError('axis must be [B,%d,3,R,C] (got %s)' % (self.SGNum, tuple(axis.shape)))
        if tuple(lamb.shape) != (B, K, R, C):
            raise ValueError('lamb must be [B,SGNum,R,C] = %s (got %s)' % ((B, K, R, C), tuple(lamb.shape)))
        if tuple(weight.shape) != (B, 3 * K, R, C):
            raise ValueError('weight must be [B,3*SGNum,R,C] = %s (got %s)' % ((B, 3 * K, R, C), tuple(weight.shape)))
        if lamb.device != axis.device or weight.device != axis.device:
            raise ValueError('axis / lamb / weight must live on one device')

    def _lazy_env(self, axis, lamb, weight, warp):
        """The env-map as a LazyTensor: evaluated by ``irb_sg_to_env_fwd`` on first use, never if the only
        consumer is ``renderingLayer.forwardEnv`` (which renders from the lobes, _lazy.py)."""
        B, K, _, R, C = axis.shape
        srcs = (axis, lamb, weight)
        vers = tuple(t._version for t in srcs)
        envH, envW = self.envHeight, self.envWidth

        def thunk():
            if tuple(t._version for t in srcs) != vers:
                raise RuntimeError('axis / lamb / weight were modified in place after output2env() and before its '
                                   'env-map was first used: the lazily evaluated env-map would not be the one the '
                                   'reference computed. Call env.materialize() right after output2env() or avoid the '
                                   'in-place edit.')
            return _SgToEnv.apply(axis, lamb, weight, envH, envW, warp)[0]
        env = LazyTensor(thunk, (B, 3, R, C, envH, envW), axis.dtype, axis.device,
                         torch.is_grad_enabled() and any(t.requires_grad for t in srcs))
        if self.SGNum <= self.MAX_FUSED_LOBES:
            env._irb_sg = (axis, lamb, weight, envH, envW, bool(warp), vers)
        return env

    def fromSGtoIm(self, axis, lamb, weight):
        """models.py:371-389: lamb / weight already warped.  Returns envmaps [B,3,R,C,envH,envW] (lazily evaluated)."""
        self._check(axis, lamb, weight)
        return self._lazy_env(axis, lamb, weight, False)

    def output2env(self, axisOrig, lambOrig, weightOrig):
        """models.py:391-404.  Returns (envmaps, axis, lamb, weight) like the reference.

        ``envmaps`` is a ``LazyTensor`` (_lazy.py): it is written by ``irb_sg_to_env_fwd`` the first time any torch
        operation touches it (wrapperBRDFLight.py:179-188 does), and never when it only travels on to
        ``renderingLayer.forwardEnv`` (testReal.py:407-411), which renders from the SG lobes it came from in the
        fused kernel.  Gradients of both consumers meet in ``axisOrig / lambOrig / weightOrig``.
        The warped ``lamb`` / ``weight`` (3rd / 4th result, read by no caller) are lazy in the same way, evaluated
        with the reference's literal expression ``tan(pi/2 * (0.999 * x))``."""
        self._check(axisOrig, lambOrig, weightOrig)
        env = self._lazy_env(axisOrig, lambOrig, weightOrig, True)

        def warped(x):
            return LazyTensor(lambda: torch.tan(np.pi / 2 * (0.999 * x)), x.shape, x.dtype, x.device,
                              torch.is_grad_enabled() and x.requires_grad)
        return env, axisOrig, warped(lambOrig), warped(weightOrig)

    def reconstLoss(self, axisOrig, lambOrig, weightOrig, envmapsBatch, segEnvBatch, offset=1.0):
        """Scale-invariant log-L2 reconstruction error of the predicted SG lighting against ``envmapsBatch``
        (the block wrapperBRDFLight.py:172-188: output2env -> LSregress on the masked env-maps -> log-L2 / pixelNum /
        3 / envWidth / envHeight) in two streaming passes over the ground truth; the predicted env-map is evaluated
        in registers and never written.  ``segEnvBatch`` is [B,1,R,C] or the reference's [B,1,R,C,1,1].

        Returns (reconstErr, coef[B]); to look at the scaled prediction call ``output2env`` and multiply by coef.
        No host synchronisation (the reference's ``pixelNum`` is a ``.item()``-style host value, here it stays on the
        device).  Extension over the reference API.
        """
        num, cnt, coef = self.reconstTerms(axisOrig, lambOrig, weightOrig, envmapsBatch, segEnvBatch, offset)
        err = num / torch.clamp(cnt, min=1e-5) / 3.0 / self.envWidth / self.envHeight
        return err, coef

    def reconstTerms(self, axisOrig, lambOrig, weightOrig, envmapsBatch, segEnvBatch, offset=1.0):
        """(numerator, sum(seg), coef[B]) of ``reconstLoss`` -- the un-normalised pair a multi-GPU caller feeds to
        ``sharding.sharded_losses`` so that the loss is normalised by the batch-GLOBAL pixel count."""
        axisOrig, lambOrig, weightOrig, envmapsBatch, segEnvBatch = [unlazy(t) for t in (axisOrig, lambOrig, weightOrig,
                                                                                          envmapsBatch, segEnvBatch)]
        self._check(axisOrig, lambOrig, weightOrig)
        B, K, _, R, C = axisOrig.shape
        _chk(envmapsBatch, 'envmapsBatch', 6)
        if tuple(envmapsBatch.shape) != (B, 3, R, C, self.envHeight, self.envWidth):
            raise ValueError('envmapsBatch must be %s (got %s)'
                             % ((B, 3, R, C, self.envHeight, self.envWidth), tuple(envmapsBatch.shape)))
        _chk(segEnvBatch, 'segEnvBatch')
        if segEnvBatch.numel() != B * R * C:
            raise ValueError('segEnvBatch must hold one value per pixel, [B,1,R,C] (got %s)' % (tuple(segEnvBatch.shape),))
        if envmapsBatch.requires_grad or segEnvBatch.requires_grad:
            raise ValueError('reconstLoss differentiates w.r.t. the SG parameters only')
        seg = segEnvBatch.reshape(B, 1, R, C)
        return _SgReconstLoss.apply(axisOrig, lambOrig, weightOrig, envmapsBatch, seg, self.envHeight, self.envWidth,
                                    float(offset))


_LAYER_TABLES = {}   # geometry -> (v, pCoord, up, ls, envWeight), see renderingLayer.__init__


class renderingLayer():
    """SG-lit microfacet shading.  Mirror of models.renderingLayer (models.py:407-522)."""

    def __init__(self, imWidth=160, imHeight=120, fov=57, F0=0.05, cameraPos=[0, 0, 0],
                 envWidth=16, envHeight=8, isCuda=True):
        self.imHeight = imHeight
        self.imWidth = imWidth
        self.envWidth = envWidth
        self.envHeight = envHeight
        self.fov = fov / 180.0 * np.pi
        self.F0 = F0
        self.cameraPos = np.array(cameraPos, dtype=np.float32).reshape([1, 3, 1, 1])
        self.xRange = 1 * np.tan(self.fov / 2)
        self.yRange = float(imHeight) / float(imWidth) * self.xRange
        self.isCuda = isCuda
        # testReal.py:341-346 builds a new layer for every image with that image's (H, W, fov): the constant tables are
        # cached per geometry (host arithmetic + 5 H2D copies happen once; afterwards construction is a dict lookup)
        key = (imWidth, imHeight, float(fov), tuple(float(x) for x in np.asarray(cameraPos).reshape(-1)), envWidth, envHeight,
               bool(isCuda), torch.cuda.current_device() if isCuda else -1)
        tabs = _LAYER_TABLES.get(key)
        if tabs is None:
            if len(_LAYER_TABLES) > 64:
                _LAYER_TABLES.clear()
            tabs = _LAYER_TABLES[key] = self._make_tables(isCuda)
        self.v, self.pCoord, self.up, self.ls, self.envWeight = tabs

    def _make_tables(self, isCuda):
        imWidth, imHeight, envWidth, envHeight = self.imWidth, self.imHeight, self.envWidth, self.envHeight
        # pixel centres on the z=-1 plane, y up (models.py:421-426); unit view vectors (:428-430)
        xs = np.linspace(-self.xRange, self.xRange, imWidth)
        ys = np.linspace(-self.yRange, self.yRange, imHeight)[::-1]
        gx, gy = np.meshgrid(xs, ys)
        pc = np.stack([gx, gy, -np.ones((imHeight, imWidth), dtype=np.float32)]).astype(np.float32)[np.newaxis]
        vv = self.cameraPos - pc
        vv = (vv / np.sqrt(np.maximum(np.sum(vv * vv, axis=1), 1e-12)[:, np.newaxis, :, :])).astype(np.float32)
        v = torch.from_numpy(vv)
        pCoord = torch.from_numpy(pc)
        up = torch.Tensor([0, 1, 0])
        lx, ly, lz, el = _direction_table(envWidth, envHeight)
        ls = np.stack([lx.reshape(-1), ly.reshape(-1), lz.reshape(-1)], axis=1)
        ew = np.sin(el.reshape(-1, 1)) * np.pi * np.pi / envWidth / envHeight
        ls = torch.from_numpy(ls.astype(np.float32))
        envWeight = torch.from_numpy(ew.astype(np.float32)).unsqueeze(0).unsqueeze(-1).unsqueeze(-1)
        if isCuda:
            v, pCoord, up, ls, envWeight = v.cuda(), pCoord.cuda(), up.cuda(), ls.cuda(), envWeight.cuda()
        return v, pCoord, up, ls, envWeight

    # -- helpers -----------------------------------------------------------------------------
    def _view(self, dev):
        v = self.v
        if v.device != dev:
            v = v.to(dev)
            self.v = v
        return v.reshape(3, self.imHeight, self.imWidth)

    def _check_maps(self, diffusePred, normalPred, roughPred, R, C):
        if any(isinstance(t, LazyTensor) for t in (diffusePred, normalPred, roughPred)):
            raise ValueError('diffusePred / normalPred / roughPred must be ordinary tensors (got a LazyTensor; pass unlazy(t))')
        _chk(diffusePred, 'diffusePred', 4)
        _chk(normalPred, 'normalPred', 4)
        _chk(roughPred, 'roughPred', 4)
        B = diffusePred.shape[0]
        if diffusePred.shape[1] != 3 or normalPred.shape[1] != 3 or roughPred.shape[1] != 1:
            raise ValueError('diffusePred / normalPred / roughPred must have 3 / 3 / 1 channels')
        if normalPred.shape[0] != B or roughPred.shape[0] != B:
            raise ValueError('batch sizes differ')
        if (R, C) != (self.imHeight, self.imWidth):
            raise ValueError('env grid %dx%d does not match the layer (imHeight=%d, imWidth=%d)'
                             % (R, C, self.imHeight, self.imWidth))

    # -- reference API -------------------------------------------------------------------------
    def forwardEnv(self, diffusePred, normalPred, roughPred, envmap):
        """models.py:461-522.  Returns (colorDiffuse, colorSpec), each [B,3,R,C].

        An env-map that came from ``output2env.output2env`` / ``fromSGtoIm`` of this package (a LazyTensor carrying its
        SG lobes) is rendered by the FUSED kernel straight from the lobes -- provided nothing modified or observed
        it in between (``LazyTensor.irb_untouched`` and the lobes' version counters); any other env tensor takes
        ``irb_render_env_fwd/bwd`` on its values."""
        if not isinstance(envmap, torch.Tensor):
            raise ValueError('envmap must be a torch.Tensor')
        if envmap.dim() != 6:
            raise ValueError('envmap must have 6 dims (got shape %s)' % (tuple(envmap.shape),))
        B, three, R, C, envH, envW = envmap.shape
        if three != 3 or (envH, envW) != (self.envHeight, self.envWidth):
            raise ValueError('envmap must be [B,3,R,C,%d,%d] (got %s)' % (self.envHeight, self.envWidth,
                                                                            tuple(envmap.shape)))
        self._check_maps(diffusePred, normalPred, roughPred, R, C)
        if isinstance(envmap, LazyTensor):
            sg = envmap._irb_sg
            if (sg is not None and envmap.irb_untouched()
                    and tuple(t._version for t in sg[:3]) == sg[6]):
                return self.forwardSG(diffusePred, normalPred, roughPred, sg[0], sg[1], sg[2], warp=sg[5])
            envmap = envmap.materialize()
        _chk(envmap, 'envmap', 6)
        dev = envmap.device
        A, N, Rg = _pool3(diffusePred, normalPred, roughPred, R, C)
        return _RenderEnv.apply(A, N, Rg, envmap, self._view(dev), self.F0)

    # -- fused entry point (superset of the reference API) -------------------------------------
    def forwardSG(self, diffusePred, normalPred, roughPred, axisOrig, lambOrig, weightOrig, warp=True):
        """output2env.output2env + forwardEnv in ONE kernel; the env-map is never materialised.

        axisOrig [B,K,3,R,C], lambOrig [B,K,R,C], weightOrig [B,3K,R,C] are the raw LightNet
        outputs (``warp=True`` applies tan(pi/2*0.999*x), models.py:396-400)."""
        _chk(axisOrig, 'axisOrig', 5)
        _chk(lambOrig, 'lambOrig', 4)
        _chk(weightOrig, 'weightOrig', 4)
        B, K, _, R, C = axisOrig.shape
        self._check_maps(diffusePred, normalPred, roughPred, R, C)
        dev = axisOrig.device
        A, N, Rg = _pool3(diffusePred, normalPred, roughPred, R, C)
        return _RenderSG.apply(A, N, Rg, axisOrig, lambOrig, weightOrig, self._view(dev), self.F0,
                               self.envHeight, self.envWidth, warp)


# ------------------------------------------------------------------------------------------
# scale-ambiguity regressions (models.py:7-84) -- PyTorch (per-sample reductions)
# ------------------------------------------------------------------------------------------

class _LSregressFused(torch.autograd.Function):
    """irb_lsregress: one streaming pass for the two sums, one for the scaling (models.py:7-21)."""

    @staticmethod
    def forward(ctx, pred, gt, origin):
        pred, gt, origin = _c(pred), _c(gt), _c(origin)
        B = pred.shape[0]
        n = pred.numel() // B
        dev = pred.device
        out = torch.empty_like(origin)
        coef = torch.empty(B, device=dev, dtype=torch.float32)
        stats = torch.empty(B, 2, device=dev, dtype=torch.float64)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().irb_lsregress(_lib.ptr(pred), _lib.ptr(gt), _lib.ptr(origin), _lib.ptr(out), _lib.ptr(coef),
                                                _lib.ptr(stats), B, n, _lib.stream_ptr(dev)), 'irb_lsregress')
        ctx.save_for_backward(coef)
        ctx.odim = origin.dim()
        return out

    @staticmethod
    def backward(ctx, g):
        coef, = ctx.saved_tensors
        return None, None, g * coef.view([-1] + [1] * (ctx.odim - 1))


def LSregress(pred, gt, origin):
    """Per-sample least-squares scale mapping pred onto gt, applied to origin (models.py:7-21).

    CUDA float32 tensors of one shape whose pred / gt carry no gradient (the reference's call pattern,
    wrapperBRDFLight.py:180-181) take the fused kernels; anything else the PyTorch expression."""
    pred, gt, origin = unlazy(pred), unlazy(gt), unlazy(origin)   # autograd.Function.apply must see real tensors
    if (pred.is_cuda and pred.dtype == gt.dtype == origin.dtype == torch.float32 and pred.shape == gt.shape == origin.shape
            and pred.dim() >= 2 and not (pred.requires_grad or gt.requires_grad)):
        return _LSregressFused.apply(pred, gt, origin)
    nb = pred.size(0)
    p = pred.reshape(nb, -1)
    g = gt.reshape(nb, -1)
    coef = (torch.sum(p * g, dim=1) / torch.clamp(torch.sum(p * p, dim=1), min=1e-5)).detach()
    coef = torch.clamp(coef, 0.001, 1000)
    return origin * coef.view([nb] + [1] * (origin.dim() - 1))


class _LSregressDiffSpecFused(torch.autograd.Function):
    """irb_lsregress_diffspec: three launches instead of ~30; coefficients are graph constants (models.py:58,76 detach)."""

    @staticmethod
    def forward(ctx, diff, spec, im, diffOrig, specOrig):
        diff, spec, im = _c(diff), _c(spec), _c(im)
        dO, sO = _c(diffOrig), _c(specOrig)
        B = diff.shape[0]
        n = diff.numel() // B
        dev = diff.device
        od, os_ = torch.empty_like(dO), torch.empty_like(sO)
        coef = torch.empty(B, 2, device=dev, dtype=torch.float32)
        stats = torch.empty(B, 8, device=dev, dtype=torch.float64)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().irb_lsregress_diffspec(
                _lib.ptr(diff), _lib.ptr(spec), _lib.ptr(im), _lib.ptr(dO), _lib.ptr(sO), _lib.ptr(od), _lib.ptr(os_),
                _lib.ptr(coef), _lib.ptr(stats), B, n, _lib.stream_ptr(dev)), 'irb_lsregress_diffspec')
        ctx.save_for_backward(coef)
        return od, os_

    @staticmethod
    def backward(ctx, g_d, g_s):
        coef, = ctx.saved_tensors
        B = coef.shape[0]
        gd = g_d * coef[:, 0].view(B, 1, 1, 1) if (g_d is not None and ctx.needs_input_grad[3]) else None
        gs = g_s * coef[:, 1].view(B, 1, 1, 1) if (g_s is not None and ctx.needs_input_grad[4]) else None
        return None, None, None, gd, gs


def LSregressDiffSpec(diff, spec, imOrig, diffOrig, specOrig):
    """Two-coefficient (diffuse, specular) per-sample regression with diffuse-only fallback, then a
    second scalar fit of clamp(d+s,0,1) to the image (models.py:23-84).

    CUDA float32 inputs whose statistics tensors (diff, spec, imOrig) carry no gradient -- the reference's
    own call pattern, wrapperBRDFLight.py:197-201 -- take the fused kernel; anything else runs the
    PyTorch expression of the same formula."""
    diff, spec, imOrig, diffOrig, specOrig = [unlazy(t) for t in (diff, spec, imOrig, diffOrig, specOrig)]
    if (diff.is_cuda and diff.dtype == torch.float32 and spec.dtype == torch.float32 and imOrig.dtype == torch.float32
            and diffOrig.dtype == torch.float32 and specOrig.dtype == torch.float32 and diff.dim() == 4
            and not (diff.requires_grad or spec.requires_grad or imOrig.requires_grad)
            and diff.shape == spec.shape == imOrig.shape == diffOrig.shape == specOrig.shape):
        return _LSregressDiffSpecFused.apply(diff, spec, imOrig, diffOrig, specOrig)
    return _LSregressDiffSpec_torch(diff, spec, imOrig, diffOrig, specOrig)


class _RenderLoss(torch.autograd.Function):
    """irb_render_loss_fwd/bwd: LSregressDiffSpec + clamp + masked L2 numerator / count (wrapperBRDFLight.py:197-207)."""

    @staticmethod
    def forward(ctx, diff, spec, im, seg):
        diff, spec, im, seg = _c(diff), _c(spec), _c(im), _c(seg)
        B, ch = diff.shape[0], diff.shape[1]
        npix = diff.numel() // (B * ch)
        dev = diff.device
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        rendered = torch.empty_like(diff)
        gbase = torch.empty_like(diff) if need else None
        coef = torch.empty(B, 2, device=dev, dtype=torch.float32)
        scratch = torch.empty(8 * B + 2, device=dev, dtype=torch.float64)
        terms = scratch[8 * B:]
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().irb_render_loss_fwd(
                _lib.ptr(diff), _lib.ptr(spec), _lib.ptr(im), _lib.ptr(seg), _lib.ptr(rendered), _lib.ptr(gbase),
                _lib.ptr(coef), _lib.ptr(scratch), _lib.ptr(terms), B, ch, npix, _lib.stream_ptr(dev)), 'irb_render_loss_fwd')
        ctx.save_for_backward(gbase, coef)
        ctx.dims = (B, ch, npix)
        t32 = terms.float()
        num, cnt = t32[0], t32[1]
        ctx.mark_non_differentiable(rendered, cnt, coef)
        return rendered, num, cnt, coef

    @staticmethod
    def backward(ctx, g_rend, g_num, g_cnt, g_coef):
        gbase, coef = ctx.saved_tensors
        B, ch, npix = ctx.dims
        if g_num is None or gbase is None:
            return None, None, None, None
        g_num = _c(g_num.detach().to(torch.float32))
        gd = torch.empty_like(gbase) if ctx.needs_input_grad[0] else None
        gs = torch.empty_like(gbase) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(gbase.device):
            _lib.check(_lib.lib().irb_render_loss_bwd(_lib.ptr(gbase), _lib.ptr(coef), _lib.ptr(g_num), _lib.ptr(gd), _lib.ptr(gs),
                                                      B, ch, npix, _lib.stream_ptr(gbase.device)), 'irb_render_loss_bwd')
        return gd, gs, None, None


def renderLossTerms(diffusePred, specularPred, imBatchSmall, segBatchSmall):
    """The render-loss block of wrapperBRDFLight.py:197-207 in three launches and without the host sync on ``pixelNum``:
    ``LSregressDiffSpec(diff.detach(), spec.detach(), im, diff, spec)`` -> ``clamp(dS + sS, 0, 1)`` -> masked L2.

    Returns ``(renderedImPred, num, cnt, coef)``: ``num = sum((rendered - im)^2 * seg)`` (differentiable w.r.t.
    ``diffusePred`` / ``specularPred``, the regression coefficients being graph constants as in models.py:58,76),
    ``cnt = sum(seg)``, ``coef[B,2]`` the total diffuse / specular scales.  ``renderErr = num / max(cnt, 1e-5) / 3``;
    a multi-GPU caller feeds ``num, cnt`` to ``sharding.sharded_losses`` so that ``cnt`` is the batch-GLOBAL count
    (wrapperBRDFLight.py:192).  ``renderedImPred`` is returned for display and carries no gradient.
    Extension over the reference API."""
    diffusePred, specularPred, imBatchSmall, segBatchSmall = [unlazy(t) for t in (diffusePred, specularPred, imBatchSmall,
                                                                                  segBatchSmall)]
    for t, n in ((diffusePred, 'diffusePred'), (specularPred, 'specularPred'), (imBatchSmall, 'imBatchSmall')):
        _chk(t, n, 4)
    _chk(segBatchSmall, 'segBatchSmall')
    if not (diffusePred.shape == specularPred.shape == imBatchSmall.shape):
        raise ValueError('diffusePred / specularPred / imBatchSmall must have one shape')
    B, ch, R, C = diffusePred.shape
    if segBatchSmall.numel() != B * R * C:
        raise ValueError('segBatchSmall must hold one value per pixel, [B,1,R,C] (got %s)' % (tuple(segBatchSmall.shape),))
    if imBatchSmall.requires_grad or segBatchSmall.requires_grad:
        raise ValueError('renderLoss differentiates w.r.t. diffusePred / specularPred only')
    return _RenderLoss.apply(diffusePred, specularPred, imBatchSmall, segBatchSmall.reshape(B, 1, R, C))


def renderLoss(diffusePred, specularPred, imBatchSmall, segBatchSmall):
    """``(renderedImPred, renderErr)`` of wrapperBRDFLight.py:197-207 (single-process normalisation)."""
    rendered, num, cnt, _ = renderLossTerms(diffusePred, specularPred, imBatchSmall, segBatchSmall)
    return rendered, num / torch.clamp(cnt, min=1e-5) / 3.0


def _LSregressDiffSpec_torch(diff, spec, imOrig, diffOrig, specOrig):
    nb, nc, nh, nw = diff.size()
    mask = (imOrig < 0.9).float()
    d = (diff * mask).view(nb, -1)
    s = (spec * mask).view(nb, -1)
    im = (imOrig * mask).view(nb, -1)
    a11 = torch.sum(d * d, dim=1)
    a22 = torch.sum(s * s, dim=1)
    a12 = torch.sum(d * s, dim=1)
    det = a11 * a22 - a12 * a12
    b1 = torch.sum(d * im, dim=1)
    b2 = torch.sum(s * im, dim=1)
    c_d = (b1 * a22 - b2 * a12) / torch.clamp(det, min=1e-2)
    c_s = (-b1 * a12 + a11 * b2) / torch.clamp(det, min=1e-2)
    c_d_only = torch.clamp(b1 / torch.clamp(a11, min=1e-5), 0.001, 1000)
    ok = ((det / (nc * nh * nw)).detach() > 1e-2).float()
    coefD = torch.clamp(ok * c_d + (1 - ok) * c_d_only, min=0, max=1000).view(nb, 1, 1, 1)
    coefS = torch.clamp(ok * c_s, min=0, max=1000).view(nb, 1, 1, 1)
    dS = coefD * diffOrig
    sS = coefS * specOrig
    r = torch.clamp(dS + sS, 0, 1).view(nb, -1)
    io = imOrig.view(nb, -1)
    coefIm = (torch.sum(r * io, dim=1) / torch.clamp(torch.sum(r * r, dim=1), min=1e-5)).detach()
    coefIm = torch.clamp(coefIm, 0.001, 1000).view(nb, 1, 1, 1)
    return coefIm * dS, coefIm * sS
