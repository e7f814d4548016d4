"""Lazily materialised tensors for the reference call pattern ``output2env.output2env -> forwardEnv``.

``models.output2env.output2env`` (models.py:391-404) returns the ``[B,3,R,C,envH,envW]`` env-map (1.89 GB at
benchmark config 2).  Two of the reference's callers only hand that tensor to ``renderingLayer.forwardEnv``
(testReal.py:407-411, trainFineTune*_cascade1.py), which here renders from the SG lobes directly -- so writing the
env-map would be pure waste.  ``LazyTensor`` is a storage-less ``torch.Tensor`` subclass that carries shape / dtype /
device / requires_grad and a thunk; *any* torch operation on it (``*``, ``.detach()``, ``.view``, ``.cpu()``,
``print`` ...) first runs the thunk once (``materialize``) and then operates on the real tensor, so callers that do
read the env-map (wrapperBRDFLight.py:179-188) see exactly the reference's values and autograd graph.

Only metadata queries (``shape``, ``size()``, ``dim()``, ``dtype``, ``device``, ``is_cuda``, ``requires_grad`` ...)
are answered without materialising.

Safety net for the fused dispatch: ``irb_untouched()`` is False as soon as the materialised tensor was
modified in place (its ``_version`` moved, which also covers writes through views), or a hook / ``retain_grad`` /
``.grad`` access observed it -- ``forwardEnv`` then takes the unfused kernel on the materialised values, i.e. the
reference's exact semantics.
"""
import torch
from torch.utils._pytree import tree_map

_T = torch.Tensor
# answered from the wrapper's own metadata, no materialisation
_META_FUNCS = {
    _T.shape.__get__, _T.dtype.__get__, _T.device.__get__, _T.is_cuda.__get__, _T.requires_grad.__get__,
    _T.ndim.__get__, _T.layout.__get__, _T.is_sparse.__get__, _T.is_quantized.__get__, _T.is_meta.__get__,
    _T.is_cpu.__get__, _T.names.__get__,
    _T.dim, _T.size, _T.numel, _T.nelement, _T.ndimension, _T.element_size, _T.is_floating_point, _T.is_complex,
    _T.get_device, _T.is_contiguous, _T.stride, _T.__len__, _T.is_leaf.__get__,
}
# these let the caller observe / intercept the gradient of THIS tensor: the fused render would bypass it
_OBSERVER_NAMES = {'register_hook', 'retain_grad', 'register_post_accumulate_grad_hook', 'requires_grad_'}
_OBSERVER_FUNCS = {_T.grad.__get__, _T.grad.__set__, _T.requires_grad.__set__, _T.grad_fn.__get__, _T.retains_grad.__get__}


def _wrapper_grad_error(grad):
    raise RuntimeError('a gradient reached a LazyTensor wrapper directly: it was passed as an input of a custom '
                       'torch.autograd.Function. Pass env.materialize() (or unlazy(env)) to such functions.')


def unlazy(t):
    """the real tensor behind a LazyTensor (materialising it), anything else unchanged"""
    return t.materialize() if isinstance(t, LazyTensor) else t


class LazyTensor(torch.Tensor):
    @staticmethod
    def __new__(cls, thunk, shape, dtype, device, requires_grad=False):
        return torch.Tensor._make_wrapper_subclass(cls, tuple(shape), dtype=dtype, device=device,
                                                   requires_grad=bool(requires_grad))

    def __init__(self, thunk, shape, dtype, device, requires_grad=False):
        self._irb_thunk = thunk
        self._irb_real = None
        self._irb_real_version = None
        self._irb_observed = False
        self._irb_grad_mode = torch.is_grad_enabled()
        self._irb_sg = None   # set by models.output2env: (axis, lamb, weight, envH, envW, warp, versions)
        if requires_grad:
            # Every torch operation reaches the REAL tensor, so autograd never has a gradient for the wrapper itself --
            # unless a custom autograd.Function took the wrapper as a direct input (Function.apply bypasses
            # __torch_function__).  That gradient would be dropped silently; fail loudly instead.
            with torch._C.DisableTorchFunctionSubclass():
                torch.Tensor.register_hook(self, _wrapper_grad_error)

    # -- state -------------------------------------------------------------------------------
    def materialize(self):
        """run the thunk once (under the grad mode of the producing call) and return the real tensor"""
        if self._irb_real is None:
            with torch.set_grad_enabled(self._irb_grad_mode):
                real = self._irb_thunk()
            self._irb_real = real
            self._irb_real_version = real._version
            self._irb_thunk = None
        return self._irb_real

    @property
    def irb_materialized(self):
        return self._irb_real is not None

    def irb_untouched(self):
        """True while nothing can have changed or observed the values since they were produced"""
        if self._irb_observed:
            return False
        return self._irb_real is None or self._irb_real._version == self._irb_real_version

    # -- dispatch ----------------------------------------------------------------------------
    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func in _META_FUNCS:
            with torch._C.DisableTorchFunctionSubclass():
                return func(*args, **kwargs)
        if func in _OBSERVER_FUNCS or getattr(func, '__name__', '') in _OBSERVER_NAMES:
            for a in args:
                if isinstance(a, LazyTensor):
                    a._irb_observed = True

        def real(a):
            return a.materialize() if isinstance(a, LazyTensor) else a
        with torch._C.DisableTorchFunctionSubclass():
            return func(*tree_map(real, args), **tree_map(real, kwargs))

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        # reached only when an operator sees the wrapper below the torch-function layer (e.g. from C++): same rule
        def real(a):
            return a.materialize() if isinstance(a, LazyTensor) else a
        return func(*tree_map(real, args), **tree_map(real, kwargs or {}))

    def __repr__(self):
        if self._irb_real is None:
            return 'LazyTensor(shape=%s, dtype=%s, device=%s, not materialised)' % (
                tuple(self.shape), self.dtype, self.device)
        return 'LazyTensor(%r)' % (self._irb_real,)
