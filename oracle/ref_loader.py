"""Import the UNMODIFIED reference modules (test / baseline infrastructure, never product).

The reference is a flat directory of Python files with no packaging, so "installing" it is a plain
copy: ``__graft_entry__.build()`` mirrors ``/root/reference/*.py`` into the git-ignored ``baseline/_ref/`` (which
travels to the GPU box with gpurun; ``/root/reference`` itself does not exist there).  This loader puts that
directory on ``sys.path`` and imports the reference's own modules under their own names:

  ``models``            -- the real render layer / CNNs (pure torch + numpy)
  ``BilateralGrid``     -- needs the two shims of SURVEY.md 8c: an empty ``pylab`` module and
                           ``cg(tol=)`` -> ``cg(rtol=)`` for SciPy >= 1.14 (+ explicit integer maxiter)
  ``BilateralLayer``, ``wrapperBRDFLight``, ``utils`` -- as they are

Used by: tests (GPU parity of the reference's own caller through compat.patch_reference, a19 shell parity),
bench.py's ``--impl reference`` / ``cpu_baseline`` leg (kind "reference") and the torch-CUDA comparator.
"""
import importlib
import os
import shutil
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = os.environ.get('IRB_REFERENCE', '/root/reference')
REF_DST = os.path.join(ROOT, 'baseline', '_ref')


def mirror():
    """copy the reference's *.py into baseline/_ref (build container only; no-op when /root/reference is absent)"""
    if not os.path.isdir(REF_SRC):
        return os.path.isdir(REF_DST)
    os.makedirs(REF_DST, exist_ok=True)
    for f in sorted(os.listdir(REF_SRC)):
        if f.endswith('.py'):
            src, dst = os.path.join(REF_SRC, f), os.path.join(REF_DST, f)
            if not os.path.exists(dst) or open(src, 'rb').read() != open(dst, 'rb').read():
                shutil.copyfile(src, dst)
    return True


def ref_dir():
    for d in (REF_DST, REF_SRC):
        if os.path.exists(os.path.join(d, 'models.py')):
            return d
    return None


def available():
    return ref_dir() is not None


def load(*names):
    """import reference modules by name, e.g. ``models, wrapper = load('models', 'wrapperBRDFLight')``"""
    d = ref_dir()
    if d is None:
        raise RuntimeError('reference not available: neither %s nor %s holds models.py' % (REF_DST, REF_SRC))
    if d not in sys.path:
        sys.path.insert(0, d)
    out = []
    for n in names:
        if n in ('BilateralGrid', 'BilateralLayer'):
            sys.modules.setdefault('pylab', types.ModuleType('pylab'))
        mod = sys.modules.get(n)
        if mod is None or os.path.dirname(os.path.abspath(getattr(mod, '__file__', ''))) != os.path.abspath(d):
            sys.modules.pop(n, None)
            mod = importlib.import_module(n)
        if n == 'BilateralGrid' and not getattr(mod, '_irb_cg_shim', False):
            import scipy.sparse.linalg as sla
            _cg = sla.cg

            def cg_shim(A, b, x0=None, M=None, maxiter=None, tol=1e-5, _cg=_cg):
                return _cg(A, b, x0=x0, M=M, maxiter=int(maxiter), rtol=tol)
            mod.cg = cg_shim
            mod._irb_cg_shim = True
        out.append(mod)
    return out[0] if len(out) == 1 else tuple(out)
