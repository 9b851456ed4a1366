"""CPU test (where /root/reference exists): oracle/paradis_assembly.py -- the restatement of the reference's model
assembly that lets BASELINE configs 1 / 4 / 5 run on the GPU box -- is pinned against the real classes:
same state_dict, strict load, bit-identical forward, parameter gradients to rounding."""
import importlib
import os
import sys

import pytest
import torch

from conftest import REFERENCE
from oracle import paradis_assembly as A
from oracle.sl_oracle import make_grids

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "model")), reason="reference tree not present")


def _purge():
    for k in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
        del sys.modules[k]


def _reference_model(cfg, lat, lon):
    sys.dont_write_bytecode = True
    _purge()
    sys.path.insert(0, REFERENCE)
    try:
        paradis = importlib.import_module("model.paradis")
        torch.manual_seed(0)
        return paradis.Paradis(A.FakeDataModule, cfg, lat, lon)
    finally:
        sys.path.remove(REFERENCE)
        _purge()


@pytest.mark.parametrize("interp", ["bilinear", "bicubic"])
@pytest.mark.parametrize("checkpointing", [False, True])
def test_assembly_restatement_is_the_reference_model(interp, checkpointing):
    cfg = A.default_cfg(interp=interp, checkpointing=checkpointing)
    lat, lon = make_grids(32, 64, True)                      # 5.625 degrees (BASELINE configs[0])
    ref = _reference_model(cfg, lat, lon)
    mine = A.Assembly(A.FakeDataModule, cfg, lat, lon, dropin=False)
    sd = ref.state_dict()
    assert list(sd) == list(mine.state_dict())
    assert all(sd[k].shape == mine.state_dict()[k].shape for k in sd)
    mine.load_state_dict(sd, strict=True)
    assert abs(ref.dt - mine.dt) < 1e-15
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 22, 32, 64, generator=g)
    y_ref, y = ref(x), mine(x)
    assert torch.equal(y_ref, y)
    w = torch.randn(y.shape, generator=g)
    (y_ref * w).sum().backward()
    (y * w).sum().backward()
    for (n1, p1), (n2, p2) in zip(ref.named_parameters(), mine.named_parameters()):
        # (broadcast-reduction gradients such as alpha_adv are summed by a threaded CPU reduction: last-bit differences)
        assert n1 == n2 and torch.allclose(p1.grad, p2.grad, rtol=1e-4, atol=1e-5 * float(p1.grad.abs().max())), n1
