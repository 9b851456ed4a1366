"""CPU tests of the drop-in boundary against the reference tree (run where /root/reference exists):
the maintainer's two-line replacement of model/padding.py and model/advection.py
leaves every parameter name and shape of the full Paradis model unchanged, so reference
checkpoints load with strict=True."""
import importlib
import os
import sys
import types

import pytest
import torch
import yaml

from conftest import REFERENCE

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "model")), reason="reference tree not present")


class AttrDict(dict):
    def __getattr__(self, k):
        v = self[k]
        return AttrDict(v) if isinstance(v, dict) else v

    def get(self, k, d=None):
        v = dict.get(self, k, d)
        return AttrDict(v) if isinstance(v, dict) else v


def _cfg():
    cfg = yaml.safe_load(open(os.path.join(REFERENCE, "config", "paradis_settings.yaml")))
    cfg["model"]["latent_size"] = 16
    cfg["model"]["velocity_vectors"] = 8
    cfg["model"]["num_layers"] = 2
    return AttrDict(cfg)


class _DM:
    class dataset:
        num_in_dyn_features = 12
        num_in_static_features = 10
    num_common_features = 5
    num_out_features = 7


def _purge():
    for k in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
        del sys.modules[k]


def _build_paradis(patched: bool):
    sys.dont_write_bytecode = True
    _purge()
    sys.path.insert(0, REFERENCE)
    try:
        if patched:   # what INTEGRATION.md section 1 asks the maintainer to do
            import paradis_model_b200.padding as pad_mod
            pkg = importlib.import_module("model")
            fake_pad = types.ModuleType("model.padding")
            fake_pad.GeoCyclicPadding = pad_mod.GeoCyclicPadding
            sys.modules["model.padding"] = fake_pad
            pkg.padding = fake_pad
            import paradis_model_b200.advection as adv_mod
            fake_adv = types.ModuleType("model.advection")
            fake_adv.NeuralSemiLagrangian = adv_mod.NeuralSemiLagrangian
            sys.modules["model.advection"] = fake_adv
            pkg.advection = fake_adv
        paradis = importlib.import_module("model.paradis")
        from oracle.sl_oracle import make_grids
        lat, lon = make_grids(32, 64, True)
        torch.manual_seed(0)
        return paradis.Paradis(_DM, _cfg(), lat, lon)
    finally:
        sys.path.remove(REFERENCE)
        _purge()


def test_full_model_state_dict_is_unchanged_by_the_drop_in():
    ref = _build_paradis(False)
    new = _build_paradis(True)
    sd_ref, sd_new = ref.state_dict(), new.state_dict()
    assert list(sd_ref) == list(sd_new)
    assert all(sd_ref[k].shape == sd_new[k].shape for k in sd_ref)
    new.load_state_dict(sd_ref, strict=True)
    adv = new.advection[0]
    assert type(adv).__module__ == "paradis_model_b200.advection"
    assert type(adv.down_projection).__name__ == "GMBlock"          # the reference's own block class
    for name in ("lat_grid", "lon_grid", "Hf", "Wf", "min_lat", "max_lat", "min_lon", "max_lon", "d_lon", "d_lat"):
        assert torch.equal(getattr(adv, name), getattr(ref.advection[0], name)), name
    assert adv.padding_interp.pad_width == ref.advection[0].padding_interp.pad_width == 2


def test_standalone_projection_names_match_reference():
    """Without the reference on the path the stand-in projections keep the same parameter names."""
    import paradis_model_b200 as P
    from oracle.sl_oracle import make_grids
    cfg = _cfg()
    lat, lon = make_grids(16, 32, True)
    mine = P.NeuralSemiLagrangian(cfg, 16, (16, 32), 8, lat, lon, "bilinear")
    sys.path.insert(0, REFERENCE)
    try:
        _purge()
        ref_adv = importlib.import_module("model.advection")
        ref = ref_adv.NeuralSemiLagrangian(cfg, 16, (16, 32), 8, lat, lon, "bilinear")
    finally:
        sys.path.remove(REFERENCE)
        _purge()
    assert sorted(mine.state_dict()) == sorted(ref.state_dict())
    assert all(mine.state_dict()[k].shape == ref.state_dict()[k].shape for k in ref.state_dict())


def test_blocks_drop_ins_have_the_reference_signatures_and_parameters():
    """paradis_model_b200.blocks.SepConv / PhysicalDownsample (optional third edit): same
    constructor arguments, parameter names and shapes as model/blocks.py:57-116, so a reference state_dict loads."""
    from paradis_model_b200 import blocks as mine
    sys.path.insert(0, REFERENCE)
    try:
        _purge()
        ref = importlib.import_module("model.blocks")
        for k in (3, 5, 7):
            a = ref.SepConv(6, 10, (16, 32), kernel_size=k, bias=True)
            b = mine.SepConv(6, 10, (16, 32), kernel_size=k, bias=True)
            assert list(a.state_dict()) == list(b.state_dict())
            assert all(a.state_dict()[n].shape == b.state_dict()[n].shape for n in a.state_dict())
            b.load_state_dict(a.state_dict(), strict=True)
            assert b.geo_padding.pad_width == a.geo_padding.pad_width == (k - 1) // 2
        a, b = ref.PhysicalDownsample(stride=4), mine.PhysicalDownsample(stride=4)
        assert list(a.state_dict()) == list(b.state_dict()) == []
        assert b.padding.pad_width == a.padding.pad_width == 2 and b.pool.stride == a.pool.stride == 4
    finally:
        sys.path.remove(REFERENCE)
        _purge()
