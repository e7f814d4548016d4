"""Install the B200-backed classes under the reference's own names.

    import models, BilateralLayer                     # the reference's modules
    from inverserenderingofindoorscene_b200 import compat
    compat.patch_reference(models, BilateralLayer)

(optionally ``compat.patch_reference(models, BilateralLayer, utils)`` to also route ``utils.predToShading``.)

After this, ``models.renderingLayer``, ``models.output2env``, ``models.LSregress``,
``models.LSregressDiffSpec``, ``BilateralLayer.BilateralFunction`` and ``BilateralLayer.BilateralLayer``
are the implementations of this package; the CNN classes stay the reference's.  ``unpatch`` restores.
"""
from . import BilateralLayer as _bl
from . import models as _models
from . import utils as _utils

_RENDER = ('renderingLayer', 'output2env', 'LSregress', 'LSregressDiffSpec')
_BILAT = ('BilateralFunction', 'BilateralLayer')
_UTILS = ('predToShading',)
_saved = {}


def patch_reference(ref_models=None, ref_bilateral_layer=None, ref_utils=None):
    for mod, names, src in ((ref_models, _RENDER, _models), (ref_bilateral_layer, _BILAT, _bl), (ref_utils, _UTILS, _utils)):
        if mod is None:
            continue
        for n in names:
            _saved.setdefault((mod, n), getattr(mod, n, None))
            setattr(mod, n, getattr(src, n))
    return ref_models, ref_bilateral_layer


def unpatch():
    for (mod, n), val in _saved.items():
        if val is not None:
            setattr(mod, n, val)
    _saved.clear()
