"""Generate the golden vectors under tests/golden/ by RUNNING THE REFERENCE ITSELF.

Run in the build container (needs /root/reference, CPU only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference has no tests and no golden vectors of its own, so parity is pinned
by executing its modules: ``model/advection.py::NeuralSemiLagrangian`` with identity
projections (so ``forward`` is exactly lines 129-169) and ``model/padding.py::GeoCyclicPadding``.
Inputs are produced by oracle.sl_oracle helpers from fixed seeds and stored next to the outputs,
so the fixtures are self-contained on the GPU box (where /root/reference does not exist).
torch version of the generating run is recorded in every file.
"""
import math
import os
import sys

import numpy as np
import torch
import yaml

sys.dont_write_bytecode = True
REF = os.environ.get("PARADIS_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from model.advection import NeuralSemiLagrangian  # noqa: E402  (reference)
from model.padding import GeoCyclicPadding  # noqa: E402  (reference)
from oracle import sl_oracle as O  # noqa: E402

DT = 21600 * 7.29212e-5 / 8  # model/paradis.py:13-14,50 with config/paradis_settings.yaml defaults


class AttrDict(dict):
    def __getattr__(self, k):
        v = self[k]
        return AttrDict(v) if isinstance(v, dict) else v


def reference_core(H, W, V, lat, lon, interpolation):
    cfg = AttrDict(yaml.safe_load(open(os.path.join(REF, "config", "paradis_settings.yaml"))))
    m = NeuralSemiLagrangian(cfg, V, (H, W), V, lat, lon, interpolation)
    m.down_projection = torch.nn.Identity()
    m.up_projection = torch.nn.Identity()
    return m


def run_case(name, H, W, B, V, poles, interpolation, kind, cells):
    lat, lon = O.make_grids(H, W, poles)
    g = torch.Generator().manual_seed(1234)
    if kind == "smooth":
        field = O.smooth_field(lat, lon, B, V, seed=3).float()
        u, v = [t.float() for t in O.smooth_velocity(lat, lon, B, V, cells, DT, seed=4)]
    else:
        field = torch.randn(B, V, H, W, generator=g)
        sig = cells * (math.pi / H) / DT
        u = torch.randn(B, V, H, W, generator=g) * sig
        v = torch.randn(B, V, H, W, generator=g) * sig
    grad_out = torch.randn(B, V, H, W, generator=g)
    f, uu, vv = [t.clone().requires_grad_(True) for t in (field, u, v)]
    m = reference_core(H, W, V, lat, lon, interpolation)
    out = m(f, uu, vv, DT)
    out.backward(grad_out)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), H=H, W=W, B=B, V=V, poles=poles, interpolation=interpolation, dt=DT,
        field=field.numpy(), u=u.numpy(), v=v.numpy(), grad_out=grad_out.numpy(), out=out.detach().numpy(),
        grad_field=f.grad.numpy(), grad_u=uu.grad.numpy(), grad_v=vv.grad.numpy(), torch_version=torch.__version__)
    print(name, "out", tuple(out.shape), float(out.abs().max()))


def run_padding():
    data = {"torch_version": torch.__version__}
    for (H, W) in [(6, 8), (9, 12)]:
        x = torch.arange(H * W, dtype=torch.float32).reshape(1, 1, H, W)
        for p in (1, 2, 3):
            data[f"pad_{H}x{W}_p{p}"] = GeoCyclicPadding(p)(x).numpy().astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "padding_index.npz"), **data)
    print("padding_index", len(data) - 1, "maps")


def run_sepconv():
    """Reference SepConv (model/blocks.py:92-116), depthwise part: padded input -> depthwise conv."""
    from model.blocks import SepConv  # reference
    data = {"torch_version": torch.__version__}
    g = torch.Generator().manual_seed(77)
    for k in (3, 5, 7):
        C, H, W = 3, 12, 16
        m = SepConv(C, 4, (H, W), kernel_size=k)
        x = torch.randn(2, C, H, W, generator=g, requires_grad=True)
        y = m.depthwise(m.geo_padding(x))
        gy = torch.randn(y.shape, generator=g)
        y.backward(gy)
        data[f"k{k}_x"] = x.detach().numpy(); data[f"k{k}_w"] = m.depthwise.weight.detach().numpy()
        data[f"k{k}_y"] = y.detach().numpy(); data[f"k{k}_gy"] = gy.numpy()
        data[f"k{k}_gx"] = x.grad.numpy(); data[f"k{k}_gw"] = m.depthwise.weight.grad.numpy()
    from model.blocks import PhysicalDownsample  # reference
    for stride in (1, 2, 4):
        x = torch.randn(2, 3, 17, 24, generator=g)
        data[f"down_s{stride}_x"] = x.numpy()
        data[f"down_s{stride}_y"] = PhysicalDownsample(stride=stride)(x).numpy()
    np.savez_compressed(os.path.join(HERE, "sepconv_depthwise.npz"), **data)
    print("sepconv_depthwise k=3,5,7")


if __name__ == "__main__":
    torch.manual_seed(0)
    run_padding()
    run_sepconv()
    for interp in ("bilinear", "bicubic"):
        tag = "bl" if interp == "bilinear" else "bc"
        run_case(f"adv_{tag}_poles_12x16_noise", 12, 16, 1, 2, True, interp, "noise", 1.5)
        run_case(f"adv_{tag}_nopoles_12x16_noise", 12, 16, 1, 2, False, interp, "noise", 1.5)
        run_case(f"adv_{tag}_poles_33x64_smooth", 33, 64, 2, 3, True, interp, "smooth", 2.0)
        run_case(f"adv_{tag}_nopoles_32x64_smooth", 32, 64, 2, 3, False, interp, "smooth", 2.0)
        run_case(f"adv_{tag}_poles_33x64_crosspole", 33, 64, 1, 2, True, interp, "noise", 6.0)
