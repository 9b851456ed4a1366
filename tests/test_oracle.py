"""CPU tests: the oracle against the golden vectors produced by the reference, against the
reference itself when it is present, and its two restatements against each other."""
import math
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, REFERENCE, golden_cases, load_golden, relmax
from oracle import sl_oracle as O


def test_padding_index_golden():
    z = np.load(os.path.join(GOLDEN, "padding_index.npz"))
    keys = [k for k in z.files if k.startswith("pad_")]
    assert len(keys) == 6
    for k in keys:
        hw, p = k.split("_")[1], int(k.split("_p")[1])
        H, W = [int(s) for s in hw.split("x")]
        row, col = O.geocyclic_source_index(H, W, p)
        assert np.array_equal(row * W + col, z[k][0, 0]), k          # bit-exact index map
        x = torch.arange(H * W, dtype=torch.float32).reshape(1, 1, H, W)
        assert np.array_equal(O.geocyclic_pad(x, p).numpy().astype(np.int32), z[k])


def test_padding_rules():
    H, W, p = 7, 10, 2
    row, col = O.geocyclic_source_index(H, W, p)
    # interior is the identity, longitude is periodic
    assert np.array_equal(row[p:-p, p:-p], np.arange(H)[:, None].repeat(W, 1))
    assert np.array_equal(col[p, :], np.mod(np.arange(-p, W + p), W))
    # caps: reflection that excludes the pole row, shifted by W/2 (model/padding.py:26-31)
    assert row[p - 1, 0] == 1 and row[0, 0] == 2 and row[-1, 0] == H - 3
    assert col[0, p] == W // 2
    with pytest.raises(AssertionError):
        O.geocyclic_source_index(4, 7, 1)


def test_padding_adjoint_is_transpose():
    torch.manual_seed(0)
    torch.manual_seed(0)
    x = torch.randn(2, 3, 6, 8, dtype=torch.float64)
    for p in (1, 2, 3):
        g = torch.randn(2, 3, 6 + 2 * p, 8 + 2 * p, dtype=torch.float64)
        lhs = (O.geocyclic_pad(x, p) * g).sum()
        rhs = (x * O.geocyclic_pad_adjoint(g, p)).sum()
        assert abs(float(lhs - rhs)) < 1e-10


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_matches_golden(name):
    """sl_advect (op-order replay) reproduces the reference's outputs bit for bit (same torch,
    same CPU); tolerance only guards against a different torch build on another box."""
    d = load_golden(name)
    lat, lon = O.make_grids(d["H"], d["W"], d["poles"])
    out, gf, gu, gv = O.sl_advect_fwd_bwd(d["field"], d["u"], d["v"], lat, lon, d["dt"], d["grad_out"],
                                          d["interpolation"])
    assert relmax(out, d["out"]) < 1e-6
    assert relmax(gu, d["grad_u"]) < 1e-5
    assert relmax(gv, d["grad_v"]) < 1e-5
    assert relmax(gf, d["grad_field"]) < 1e-5
    if str(d["torch_version"]) == torch.__version__:
        assert torch.equal(out, d["out"])
        assert torch.equal(gu, d["grad_u"]) and torch.equal(gv, d["grad_v"])


@pytest.mark.parametrize("interp", ["bilinear", "bicubic"])
@pytest.mark.parametrize("poles", [True, False])
def test_explicit_closed_form_matches_autograd_fp64(interp, poles):
    H, W, B, V, dt = 17, 24, 2, 2, 0.19688724
    lat, lon = O.make_grids(H, W, poles, torch.float64)
    g = torch.Generator().manual_seed(5)
    field = torch.randn(B, V, H, W, generator=g, dtype=torch.float64)
    sig = 3.0 * math.pi / H / dt
    u = torch.randn(B, V, H, W, generator=g, dtype=torch.float64) * sig
    v = torch.randn(B, V, H, W, generator=g, dtype=torch.float64) * sig
    go = torch.randn(B, V, H, W, generator=g, dtype=torch.float64)
    a = O.sl_advect_fwd_bwd(field, u, v, lat, lon, dt, go, interp)
    b = O.sl_advect_explicit(field, u, v, lat, lon, dt, go, interp)
    for x, y in zip(a, b):
        assert relmax(y, x) < 1e-11


def test_zero_velocity_is_pole_fixed_identity():
    H, W = 16, 32
    for poles in (True, False):
        lat, lon = O.make_grids(H, W, poles)
        f = torch.randn(1, 2, H, W, generator=torch.Generator().manual_seed(H))
        z = torch.zeros(1, 2, H, W)
        out = O.sl_advect(f, z, z, lat, lon, 0.2, "bilinear")
        # on a pole-including grid the asin clamp (advection.py:90) keeps the pole rows
        # 4.5e-4 rad short of the pole, so identity holds only to ~1e-3 there
        assert relmax(out, O.pole_mean(f)) < (2e-3 if poles else 2e-5)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "model")), reason="reference tree not present")
@pytest.mark.parametrize("interp", ["bilinear", "bicubic"])
def test_oracle_bitexact_vs_reference_module(interp):
    sys.dont_write_bytecode = True
    sys.path.insert(0, os.path.join(GOLDEN))
    sys.path.insert(0, REFERENCE)
    try:
        import make_golden as MG
        H, W, B, V, dt = 31, 48, 2, 2, MG.DT
        lat, lon = O.make_grids(H, W, True)
        g = torch.Generator().manual_seed(9)
        field = torch.randn(B, V, H, W, generator=g)
        u = torch.randn(B, V, H, W, generator=g) * 0.4
        v = torch.randn(B, V, H, W, generator=g) * 0.4
        go = torch.randn(B, V, H, W, generator=g)
        m = MG.reference_core(H, W, V, lat, lon, interp)
        f, uu, vv = [t.clone().requires_grad_(True) for t in (field, u, v)]
        out = m(f, uu, vv, dt)
        out.backward(go)
        o2, gf, gu, gv = O.sl_advect_fwd_bwd(field, u, v, lat, lon, dt, go, interp)
        assert torch.equal(out, o2) and torch.equal(uu.grad, gu) and torch.equal(vv.grad, gv)
        assert relmax(gf, f.grad) < 1e-6
    finally:
        sys.path.remove(REFERENCE)
        sys.path.remove(GOLDEN)


def test_depthwise_oracle_matches_reference_sepconv_golden():
    z = np.load(os.path.join(GOLDEN, "sepconv_depthwise.npz"))
    for k in (3, 5, 7):
        x = torch.from_numpy(z[f"k{k}_x"]).requires_grad_(True)
        w = torch.from_numpy(z[f"k{k}_w"]).requires_grad_(True)
        y = O.geocyclic_depthwise(x, w)
        y.backward(torch.from_numpy(z[f"k{k}_gy"]))
        assert relmax(y.detach(), torch.from_numpy(z[f"k{k}_y"])) < 1e-6
        assert relmax(x.grad, torch.from_numpy(z[f"k{k}_gx"])) < 1e-5
        assert relmax(w.grad, torch.from_numpy(z[f"k{k}_gw"])) < 1e-5
