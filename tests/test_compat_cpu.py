import types

from inverserenderingofindoorscene_b200 import BilateralLayer, compat, models


def test_patch_and_unpatch():
    ref_m = types.ModuleType('models')
    ref_b = types.ModuleType('BilateralLayer')
    sentinel = object()
    ref_m.renderingLayer = sentinel
    ref_m.encoder0 = sentinel
    compat.patch_reference(ref_m, ref_b)
    assert ref_m.renderingLayer is models.renderingLayer and ref_m.output2env is models.output2env
    assert ref_m.LSregressDiffSpec is models.LSregressDiffSpec and ref_m.encoder0 is sentinel
    assert ref_b.BilateralFunction is BilateralLayer.BilateralFunction
    assert ref_b.BilateralLayer is BilateralLayer.BilateralLayer
    compat.unpatch()
    assert ref_m.renderingLayer is sentinel


def test_patch_utils():
    from inverserenderingofindoorscene_b200 import utils
    ref_u = types.ModuleType('utils')
    keep = object()
    ref_u.predToShading = keep
    ref_u.writeEnvToFile = keep
    compat.patch_reference(None, None, ref_u)
    assert ref_u.predToShading is utils.predToShading and ref_u.writeEnvToFile is keep
    compat.unpatch()
    assert ref_u.predToShading is keep


def test_pred_to_shading_refuses_cpu():
    import numpy as np
    import pytest
    import torch
    from inverserenderingofindoorscene_b200 import utils
    with pytest.raises(ValueError):
        utils.predToShading(torch.zeros(1, 84, 4, 4))          # CPU tensor
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            utils.predToShading(np.zeros((1, 84, 4, 4), np.float32))


def test_bilateral_layer_shell_cpu():
    import torch
    layer = BilateralLayer.BilateralLayer(mode=2, isCuda=False)
    keys = sorted(set(k.split('.')[0] for k in layer.state_dict()))
    assert keys == sorted(['conv1', 'gn1', 'conv2', 'gn2', 'dconv1', 'dgn1', 'dconv2', 'dgn2', 'dconvFinal'])
    assert sum(p.numel() for p in layer.parameters()) > 10000
    assert layer.conv1.in_channels == 4
    assert [round(float(x), 6) for x in layer.grid_params_arr] == [8, 2, 8]
    assert [float(x) for x in layer.bs_params_arr][0] == 300 and int(layer.bs_params_arr[3]) == 10
    assert layer.computePadding(5, 8) == [2, 1] and layer.computePadding(4, 8) == [2, 2]


def test_cascade_handoff_formats(tmp_path, monkeypatch):
    """SG volume packing (wrapperBRDFLight.py:168-169) and the HDF5 hand-off calls (utils.py:92-99, dataLoader.py:277-283)
    against a recording stand-in for h5py (not installed in this image): dataset name, compression and array layout."""
    import sys
    import types

    import numpy as np
    import torch

    from inverserenderingofindoorscene_b200 import utils as u
    B, K, R, C = 2, 12, 3, 4
    axis, lamb, wt = torch.rand(B, K, 3, R, C), torch.rand(B, K, R, C), torch.rand(B, 3 * K, R, C)
    vol = u.packSGVolume(axis, lamb, wt)
    assert tuple(vol.shape) == (B, 7 * K, R, C)
    assert torch.equal(vol[:, :36], axis.view(B, 36, R, C)) and torch.equal(vol[:, 36:48], lamb) and torch.equal(vol[:, 48:], wt)
    a2, l2, w2 = u.unpackSGVolume(vol, SGNum=K)
    assert torch.equal(a2, axis) and torch.equal(l2, lamb) and torch.equal(w2, wt)
    store = {}

    class _File(object):
        def __init__(self, name, mode):
            self.name, self.mode = str(name), mode

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

        def create_dataset(self, key, data=None, compression=None):
            store[self.name] = (key, np.array(data), compression)

        def get(self, key):
            k, d, _ = store[self.name]
            return d if k == key else None
    monkeypatch.setitem(sys.modules, 'h5py', types.SimpleNamespace(File=_File))
    names = [str(tmp_path / ('imdiffuse_%d_0.h5' % i)) for i in range(B)]
    diff = torch.rand(B, 3, R, C)
    u.writeH5ToFile(diff, names)
    for i, n in enumerate(names):
        key, data, comp = store[n]
        assert key == 'data' and comp == 'lzf' and data.shape == (3, R, C) and np.array_equal(data, diff[i].numpy())
        assert np.array_equal(u.loadH5(n), diff[i].numpy())
    assert u.loadH5(str(tmp_path / 'missing.h5')) is None
