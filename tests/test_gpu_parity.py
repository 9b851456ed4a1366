"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C ABI
(libparadis_sl.so via ctypes / torch.library); the oracle is only the checker."""
import math

import numpy as np
import pytest
import torch

from conftest import bad_fraction, golden_cases, load_golden, relmax
from oracle import sl_oracle as O

pytestmark = pytest.mark.gpu
DT = 21600 * 7.29212e-5 / 8


@pytest.fixture(autouse=True, params=["default-backward", "row-sweep-forced"])
def _backward_kernel(request):
    """Every test runs twice: with the library's own choice of fused backward (row sweep on wide bilinear meshes, strip
    sweep otherwise) and with the warp-specialised row sweep forced wherever its plan fits (PARADIS_BWD_ROWSWEEP)."""
    from paradis_model_b200 import ops
    ops.FORCE_ROW_SWEEP = request.param == "row-sweep-forced"
    yield
    ops.FORCE_ROW_SWEEP = False


def P():
    import paradis_model_b200 as pkg
    return pkg


def cuda_fwd_bwd(field, u, v, lat, lon, dt, go, interp, math_mode="fast", pole_fix=True, cfl=8.0):
    pkg = P()
    geo = pkg.SLGeometry.from_grids(lat.cuda(), lon.cuda())
    f, uu, vv = [t.cuda().clone().requires_grad_(True) for t in (field, u, v)]
    out = pkg.sl_advect(f, uu, vv, geo, dt, interp, pole_fix, math_mode, cfl)
    out.backward(go.cuda())
    pkg.check_status()
    return out.detach().cpu(), f.grad.cpu(), uu.grad.cpu(), vv.grad.cpu()


# ------------------------------------------------------------------ padding
@pytest.mark.parametrize("p", [1, 2, 3])
@pytest.mark.parametrize("hw", [(6, 8), (32, 64), (721, 1440)])
def test_padding_bit_exact(hw, p):
    H, W = hw
    x = torch.arange(H * W, dtype=torch.float32).reshape(1, 1, H, W).repeat(1, 2, 1, 1)
    x[:, 1] += 0.5
    y = P().geocyclic_pad(x.cuda(), p).cpu()
    assert torch.equal(y, O.geocyclic_pad(x, p))            # index map bit-exact


@pytest.mark.parametrize("p", [1, 2, 3])
def test_padding_backward(p):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 10, 16, generator=g)
    gy = torch.randn(2, 3, 10 + 2 * p, 16 + 2 * p, generator=g)
    xc = x.cuda().requires_grad_(True)
    P().geocyclic_pad(xc, p).backward(gy.cuda())
    assert relmax(xc.grad.cpu(), O.geocyclic_pad_adjoint(gy, p)) < 1e-6
    gi = torch.randint(-8, 8, gy.shape, generator=g).float()      # integers: order-independent, exact
    xc.grad = None
    P().geocyclic_pad(xc, p).backward(gi.cuda())
    assert torch.equal(xc.grad.cpu(), O.geocyclic_pad_adjoint(gi, p))


def test_padding_module_asserts():
    m = P().GeoCyclicPadding(1)
    with pytest.raises(AssertionError):
        m(torch.zeros(1, 1, 4, 7, device="cuda"))
    with pytest.raises(AssertionError):
        m(torch.zeros(1, 4, 8, device="cuda"))
    x = torch.zeros(1, 1, 4, 8, device="cuda")
    assert P().GeoCyclicPadding(0)(x) is x


# ------------------------------------------------------------------ golden vectors
@pytest.mark.parametrize("math_mode", ["fast", "exact"])
@pytest.mark.parametrize("name", golden_cases())
def test_golden(name, math_mode):
    d = load_golden(name)
    lat, lon = O.make_grids(d["H"], d["W"], d["poles"])
    out, gf, gu, gv = cuda_fwd_bwd(d["field"], d["u"], d["v"], lat, lon, d["dt"], d["grad_out"],
                                   d["interpolation"], math_mode)
    smooth = "smooth" in name
    assert relmax(out, d["out"]) < (1e-5 if smooth else 1e-4), "forward"
    assert relmax(gf, d["grad_field"]) < 1e-4, "grad_field"
    if smooth:   # north_star tolerances: fwd 1e-5, grads 1e-4 (relative to max)
        assert relmax(gu, d["grad_u"]) < 1e-4 and relmax(gv, d["grad_v"]) < 1e-4
    else:        # white noise: d out / d ix jumps across cell edges, allow isolated floor flips
        assert bad_fraction(gu, d["grad_u"], 1e-4) < 2e-3 and bad_fraction(gv, d["grad_v"], 1e-4) < 2e-3


# ------------------------------------------------------------------ oracle at model sizes
@pytest.mark.parametrize("math_mode", ["fast", "exact"])
@pytest.mark.parametrize("interp", ["bilinear", "bicubic"])
@pytest.mark.parametrize("poles,H,W", [(False, 128, 256), (True, 181, 360)])
def test_smooth_fields_vs_oracle(poles, H, W, interp, math_mode):
    B, V = 2, 4
    lat, lon = O.make_grids(H, W, poles)
    field = O.smooth_field(lat, lon, B, V).float()
    u, v = [t.float() for t in O.smooth_velocity(lat, lon, B, V, 2.5, DT)]
    go = O.smooth_field(lat, lon, B, V, seed=7).float()
    ref = O.sl_advect_fwd_bwd(field, u, v, lat, lon, DT, go, interp)
    got = cuda_fwd_bwd(field, u, v, lat, lon, DT, go, interp, math_mode)
    assert relmax(got[0], ref[0]) < 1e-5, "forward"
    assert relmax(got[1], ref[1]) < 1e-4, "grad_field"
    # d out / d ix is piecewise constant in ix (bilinear) so a departure point within an ulp of a
    # cell edge may pick the other cell than the CPU oracle; on this coarse mesh that is a ~5e-3
    # jump at an isolated point (the reference's own fp32-vs-fp64 noise shows the same, see
    # tools/diag_poles.py).  Everything else must meet the north_star 1e-4.
    for a, b, name in ((got[2], ref[2], "grad_u"), (got[3], ref[3], "grad_v")):
        assert bad_fraction(a, b, 1e-4) < 1e-4, name
        assert relmax(a, b) < 3e-2, name


@pytest.mark.parametrize("interp", ["bilinear", "bicubic"])
def test_exact_mode_vs_same_device_torch(interp):
    """White noise, against the oracle's op-order replay run by torch on the SAME GPU: the
    departure coordinates are bit-identical, so even white noise agrees to rounding."""
    H, W, B, V = 128, 256, 2, 8
    lat, lon, field, u, v, go = O.bench_inputs(H, W, B, V, False, DT)
    dev = [t.cuda() for t in (field, u, v, lat, lon, go)]
    ref = O.sl_advect_fwd_bwd(dev[0], dev[1], dev[2], dev[3], dev[4], DT, dev[5], interp)
    got = cuda_fwd_bwd(field, u, v, lat, lon, DT, go, interp, "exact")
    geo = O.Geometry(dev[3], dev[4])
    errs = [relmax(a, b.cpu()) for a, b in zip(got, ref)]
    bad = [bad_fraction(a, b.cpu(), 1e-4) for a, b in zip(got, ref)]
    print("exact-vs-torch-cuda", interp, errs, bad)
    assert errs[0] < 1e-4 and bad[0] == 0.0, errs      # white-noise field: one ulp of ix is ~3e-5 of max|out|
    assert errs[1] < 1e-4, errs
    # identical coordinates => no cell-edge flips: grad_u / grad_v must meet the north_star 1e-4 EVERYWHERE
    assert errs[2] < 1e-4 and errs[3] < 1e-4 and bad[2] == 0.0 and bad[3] == 0.0, (errs, bad)


# ------------------------------------------------------------------ value parity at the headline size (0.25 deg)
_C3_ORACLE = {}


def _c3_smooth(interp):
    """721x1440 pole-including mesh, smooth field and ~4-cell smooth velocities; the oracle run by torch on the CPU
    and on this GPU (cached per stencil), and the difference between the two: the reference's own CPU-vs-CUDA noise."""
    if interp not in _C3_ORACLE:
        H, W, B, V = 721, 1440, 1, 3
        lat, lon = O.make_grids(H, W, True)
        field = O.smooth_field(lat, lon, B, V).float()
        u, v = [t.float() for t in O.smooth_velocity(lat, lon, B, V, 4.0, DT)]
        go = O.smooth_field(lat, lon, B, V, seed=7).float()
        ref_cpu = O.sl_advect_fwd_bwd(field, u, v, lat, lon, DT, go, interp)
        dev = [t.cuda() for t in (field, u, v, lat, lon, go)]
        ref_gpu = [t.cpu() for t in O.sl_advect_fwd_bwd(dev[0], dev[1], dev[2], dev[3], dev[4], DT, dev[5], interp)]
        noise = [relmax(a, b) for a, b in zip(ref_gpu, ref_cpu)]
        _C3_ORACLE[interp] = (lat, lon, field, u, v, go, ref_cpu, ref_gpu, noise)
    return _C3_ORACLE[interp]


@pytest.mark.parametrize("cfl", [6.0, 0.0], ids=["rowsweep", "general"])
@pytest.mark.parametrize("math_mode", ["fast", "exact"])
@pytest.mark.parametrize("interp", ["bilinear", "bicubic"])
def test_c3_size_values_vs_cpu_oracle(interp, math_mode, cfl):
    """out, grad_field, grad_u, grad_v at 721x1440 (polar rows with 1/cos(lat) reach, cap folds, pole means, both
    backward paths) against the oracle run by torch on the CPU and on this GPU.  north_star: forward 1e-5, gradients
    1e-4 (relative to max).
    * forward: 1e-5 against both, every mode;
    * EXACT mode (the reference's fp32 operation order): gradients 1e-4 EVERYWHERE against torch on this GPU;
    * FAST mode, and anything against the CPU run: the yardstick is the reference itself -- its CPU and CUDA runs differ
      from each other by 3e-4 (grad_field), 7e-3 (grad_u), 1.4e-2 (grad_v) of the maximum at this size (one ulp of
      longitude is 1.1e-4 cell at 0.25 degrees; d out / d ix is piecewise constant; d lat / d sin(lat) reaches 2e3 at the
      poles), measured here as `noise`.  The gradients must be within 1e-4 or three times that, whichever is larger,
      and within 1e-4 on all but 2e-3 of the points."""
    lat, lon, field, u, v, go, ref_cpu, ref_gpu, noise = _c3_smooth(interp)
    got = cuda_fwd_bwd(field, u, v, lat, lon, DT, go, interp, math_mode, cfl=cfl)
    e_cpu = [relmax(a, b) for a, b in zip(got, ref_cpu)]
    e_gpu = [relmax(a, b) for a, b in zip(got, ref_gpu)]
    print("c3 smooth", interp, math_mode, cfl, "vs cpu", e_cpu, "vs same-gpu torch", e_gpu, "reference cpu-vs-cuda", noise)
    assert e_gpu[0] < 1e-5 and e_cpu[0] < 1e-5, (e_gpu, e_cpu)
    if math_mode == "exact":
        assert e_gpu[1] < 1e-4 and e_gpu[2] < 1e-4 and e_gpu[3] < 1e-4, e_gpu
    for k in (1, 2, 3):
        bound = max(1e-4, 3 * noise[k])
        assert e_cpu[k] < bound and e_gpu[k] < bound, (k, e_cpu, e_gpu, noise)
        assert bad_fraction(got[k], ref_gpu[k], 1e-4) < 2e-3 or e_gpu[k] < 5e-4, k


@pytest.mark.parametrize("interp", ["bilinear", "bicubic"])
def test_c3_size_exact_mode_white_noise_vs_same_device_torch(interp):
    """White noise at 721x1440, V=8, EXACT mode, row-sweep backward, against the oracle's op replay run by torch on
    the same GPU (bit-identical coordinates): every output within tolerance, no outliers."""
    H, W, B, V = 721, 1440, 1, 8
    lat, lon, field, u, v, go = O.bench_inputs(H, W, B, V, True, DT)
    dev = [t.cuda() for t in (field, u, v, lat, lon, go)]
    ref = [t.cpu() for t in O.sl_advect_fwd_bwd(dev[0], dev[1], dev[2], dev[3], dev[4], DT, dev[5], interp)]
    got = cuda_fwd_bwd(field, u, v, lat, lon, DT, go, interp, "exact", cfl=6.0)
    errs = [relmax(a, b) for a, b in zip(got, ref)]
    print("c3 exact-vs-torch-cuda", interp, errs)
    assert errs[0] < 1e-5, errs                      # identical stencil indices and weights: rounding of the sums only
    assert errs[1] < 1e-4 and errs[2] < 1e-4 and errs[3] < 1e-4, errs


def test_fast_mode_white_noise_statistics():
    """fast math vs CPU oracle on white noise: isolated floor flips only."""
    H, W, B, V = 128, 256, 2, 8
    lat, lon, field, u, v, go = O.bench_inputs(H, W, B, V, False, DT)
    ref = O.sl_advect_fwd_bwd(field, u, v, lat, lon, DT, go, "bilinear")
    got = cuda_fwd_bwd(field, u, v, lat, lon, DT, go, "bilinear", "fast")
    assert bad_fraction(got[0], ref[0], 1e-4) < 1e-3
    assert bad_fraction(got[1], ref[1], 1e-4) < 1e-3
    assert bad_fraction(got[2], ref[2], 1e-3) < 2e-3


# ------------------------------------------------------------------ structure
@pytest.mark.parametrize("interp", ["bilinear", "bicubic"])
def test_backward_is_deterministic(interp):
    H, W, B, V = 96, 192, 2, 6
    lat, lon, field, u, v, go = O.bench_inputs(H, W, B, V, True, DT, cells_sigma=3.0, cells_clip=8.0)
    a = cuda_fwd_bwd(field, u, v, lat, lon, DT, go, interp)
    for _ in range(3):
        b = cuda_fwd_bwd(field, u, v, lat, lon, DT, go, interp)
        for x, y in zip(a, b):
            assert torch.equal(x, y)


def test_zero_velocity_identity():
    H, W = 64, 128
    lat, lon = O.make_grids(H, W, False)
    f = torch.randn(1, 3, H, W, generator=torch.Generator().manual_seed(11))   # seeded: an unseeded draw once landed at 1.02e-5
    z = torch.zeros(1, 3, H, W)
    geo = P().SLGeometry.from_grids(lat.cuda(), lon.cuda())
    out = P().sl_advect(f.cuda(), z.cuda(), z.cuda(), geo, DT, "bilinear").cpu()
    assert relmax(out, O.pole_mean(f)) < 1e-4      # white-noise field, coordinates good to ~1e-5 cells
    out = P().sl_advect(f.cuda(), z.cuda(), z.cuda(), geo, DT, "bilinear", True, "exact").cpu()
    assert relmax(out, O.pole_mean(f)) < 5e-5      # the reference's own normalise/un-normalise round trip
    assert relmax(out, O.sl_advect(f, z, z, lat, lon, DT, "bilinear")) < 2e-5   # white-noise field x ~1e-5 cells of coordinate rounding


def test_strided_velocity_views_and_no_pole_fix():
    """u, v as views of one [B, 2V, H, W] tensor (model/paradis.py:235-237)."""
    H, W, B, V = 48, 96, 3, 5
    lat, lon = O.make_grids(H, W, True)
    g = torch.Generator().manual_seed(2)
    vel = (torch.randn(B, 2 * V, H, W, generator=g) * 0.05).cuda()
    velv = vel.view(B, 2, V, H, W)
    u, v = velv[:, 0], velv[:, 1]
    assert not u.is_contiguous()
    field = torch.randn(B, V, H, W, generator=g)
    geo = P().SLGeometry.from_grids(lat.cuda(), lon.cuda())
    for pole_fix in (True, False):
        out = P().sl_advect(field.cuda(), u, v, geo, DT, "bilinear", pole_fix).cpu()
        ref = O.sl_advect(field, u.cpu().contiguous(), v.cpu().contiguous(), lat, lon, DT, "bilinear", pole_fix)
        assert bad_fraction(out, ref, 1e-4) < 1e-3


@pytest.mark.parametrize("interp", ["bilinear", "bicubic"])
def test_adjoint_identity_full_size(interp):
    """Size-independent property at the BASELINE size (721x1440): the operator is linear in
    `field`, so <A f, g> == <f, A^T g>.  Checks grad_field (inverse-stencil gather, pole folds,
    longitude wrap) against the forward kernel without any oracle."""
    H, W, B, V = 721, 1440, 1, 8
    lat, lon, field, u, v, go = O.bench_inputs(H, W, B, V, True, DT)
    pkg = P()
    geo = pkg.SLGeometry.from_grids(lat.cuda(), lon.cuda())
    f = field.cuda().requires_grad_(True)
    out = pkg.sl_advect(f, u.cuda(), v.cuda(), geo, DT, interp)
    out.backward(go.cuda())
    lhs = (out.detach().double() * go.cuda().double()).sum()
    rhs = (f.grad.double() * field.cuda().double()).sum()
    pkg.check_status()
    assert abs(float(lhs - rhs)) / abs(float(lhs)) < 1e-5
    # linearity: A(2f + h) == 2 A f + A h
    h = torch.randn_like(field).cuda()
    o2 = pkg.sl_advect(2 * field.cuda() + h, u.cuda(), v.cuda(), geo, DT, interp)
    o3 = pkg.sl_advect(h, u.cuda(), v.cuda(), geo, DT, interp)
    assert relmax((2 * out.detach() + o3).cpu(), o2.cpu()) < 1e-5


def test_large_displacement_window():
    """Displacements of ~20 rows: the gather window follows the measured reach."""
    H, W, B, V = 91, 180, 1, 3
    lat, lon, field, u, v, go = O.bench_inputs(H, W, B, V, True, DT, cells_sigma=8.0, cells_clip=24.0)
    ref = O.sl_advect_fwd_bwd(field, u, v, lat, lon, DT, go, "bilinear")
    got = cuda_fwd_bwd(field, u, v, lat, lon, DT, go, "bilinear", "fast")
    assert bad_fraction(got[0], ref[0], 1e-4) < 2e-3
    assert bad_fraction(got[1], ref[1], 1e-4) < 2e-3


def test_lat_band_windows_are_bit_identical():
    """Two latitude bands with halos reproduce the single-call result bit for bit."""
    H, W, B, V = 64, 128, 1, 4
    lat, lon, field, u, v, go = O.bench_inputs(H, W, B, V, True, DT, cells_sigma=1.0, cells_clip=2.0)
    pkg = P()
    geo = pkg.SLGeometry.from_grids(lat.cuda(), lon.cuda())
    fc, uc, vc, gc = [t.cuda() for t in (field, u, v, go)]
    full = torch.ops.paradis.sl_advect(fc, uc, vc, geo.tables, geo.scalars, DT, 1, True, 0, geo.windows, 0.0)
    gfull = torch.ops.paradis.sl_advect_backward(gc, fc, uc, vc, geo.tables, geo.scalars, DT, 1, True, 0,
                                                 geo.windows, 0.0, True, True)
    halo = 6
    outs, gfs, gus = [], [], []
    for (r0, n) in [(0, 32), (32, 32)]:
        f0, f1 = max(0, r0 - halo), min(H, r0 + n + halo)
        gb = geo.band((r0, n), (r0, n), (f0, f1 - f0))
        outs.append(torch.ops.paradis.sl_advect(fc[:, :, f0:f1].contiguous(), uc[:, :, r0:r0 + n].contiguous(),
                                                vc[:, :, r0:r0 + n].contiguous(), gb.tables, gb.scalars, DT, 1,
                                                True, 0, gb.windows, 0.0))
        gb2 = geo.band((r0, n), (f0, f1 - f0), (f0, f1 - f0))
        r = torch.ops.paradis.sl_advect_backward(gc[:, :, f0:f1].contiguous(), fc[:, :, f0:f1].contiguous(),
                                                 uc[:, :, f0:f1].contiguous(), vc[:, :, f0:f1].contiguous(),
                                                 gb2.tables, gb2.scalars, DT, 1, True, 0, gb2.windows, 0.0, True,
                                                 True)
        gfs.append(r[0]); gus.append(r[1])
    pkg.check_status()
    assert torch.equal(torch.cat(outs, 2), full)
    assert torch.equal(torch.cat(gfs, 2), gfull[0])
    assert torch.equal(torch.cat(gus, 2), gfull[1])


def test_halo_violation_is_reported():
    H, W = 64, 128
    lat, lon, field, u, v, go = O.bench_inputs(H, W, 1, 1, True, DT, cells_sigma=4.0, cells_clip=8.0)
    pkg = P()
    geo = pkg.SLGeometry.from_grids(lat.cuda(), lon.cuda()).band((20, 10), (20, 10), (19, 12))
    torch.ops.paradis.sl_advect(field[:, :, 19:31].contiguous().cuda(), u[:, :, 20:30].contiguous().cuda(),
                                v[:, :, 20:30].contiguous().cuda(), geo.tables, geo.scalars, DT, 1, True, 0,
                                geo.windows, 0.0)
    with pytest.raises(RuntimeError, match="DISPLACEMENT"):
        pkg.check_status()


def test_host_entry_matches_device_path():
    from paradis_model_b200 import ops
    ops.FORCE_ROW_SWEEP = False            # the host entry makes the library's own kernel choice: compare like with like
    H, W, B, V = 64, 128, 2, 5
    lat, lon, field, u, v, go = O.bench_inputs(H, W, B, V, False, DT)
    pkg = P()
    geo = pkg.SLGeometry.from_grids(lat.cuda(), lon.cuda())
    dev = cuda_fwd_bwd(field, u, v, lat, lon, DT, go, "bilinear")
    pin = lambda t: t.contiguous().pin_memory()
    hin = [pin(t) for t in (field, u, v, go)]
    hout = [pin(torch.empty_like(field)) for _ in range(4)]
    pkg.host_fwd_bwd(geo, *hin, *hout, DT, "bilinear", True, "fast", chunk_planes=3)
    for a, b in zip(hout, dev):
        assert torch.equal(a, b)


def test_module_drop_in_autograd_and_state_dict():
    """NeuralSemiLagrangian keeps the reference's parameter names and differentiates end to end."""
    pkg = P()

    class NS(dict):
        __getattr__ = dict.__getitem__

    cfg = NS(model=NS(physblock=NS(advection=NS(down_projection=NS(layers=["SepConv"], hidden_dim=0),
                                                up_projection=NS(layers=["CLinear"], hidden_dim=0)))))
    H, W, hidden, V, B = 32, 64, 12, 6, 2
    lat, lon = O.make_grids(H, W, True)
    torch.manual_seed(0)
    m = pkg.NeuralSemiLagrangian(cfg, hidden, (H, W), V, lat, lon, "bilinear").cuda()
    assert sorted(m.state_dict()) == ["down_projection.0-SepConv.depthwise.weight",
                                      "down_projection.0-SepConv.pointwise.bias",
                                      "down_projection.0-SepConv.pointwise.weight",
                                      "up_projection.0-CLinear.conv.bias", "up_projection.0-CLinear.conv.weight"]
    hid = torch.randn(B, hidden, H, W, device="cuda", requires_grad=True)
    vel = (torch.randn(B, 2 * V, H, W, device="cuda") * 0.05).requires_grad_(True)
    velv = vel.view(B, 2, V, H, W)
    out = m(hid, velv[:, 0], velv[:, 1], DT)
    out.square().mean().backward()
    assert out.shape == (B, hidden, H, W)
    assert hid.grad is not None and vel.grad is not None and torch.isfinite(vel.grad).all()
    # same computation with the oracle core in place of the fused op
    proj = m.down_projection(hid.detach()).cpu()
    core = O.sl_advect(proj, velv[:, 0].detach().cpu(), velv[:, 1].detach().cpu(), lat, lon, DT, "bilinear")
    ref = m.up_projection(core.cuda())
    # white-noise field and velocities through a random 1x1 projection: fast-math coordinates differ from
    # the CPU oracle by ~1e-5 cell, i.e. ~1e-4 of max|out| at the steepest points
    assert bad_fraction(out.detach().cpu(), ref.detach().cpu(), 3e-4) < 1e-3


# ------------------------------------------------------------------ fused backward sweep
@pytest.mark.parametrize("interp", ["bilinear", "bicubic"])
@pytest.mark.parametrize("poles,H,W", [(False, 128, 256), (True, 181, 360)])
def test_sweep_matches_general_path(poles, H, W, interp):
    """The fused single-pass backward (cfl_cells > 0) and the general two-kernel path compute the
    same sums in a different (but each fixed) order."""
    B, V = 2, 5
    lat, lon, field, u, v, go = O.bench_inputs(H, W, B, V, poles, DT)       # +-4 cells
    gen = cuda_fwd_bwd(field, u, v, lat, lon, DT, go, interp, cfl=0.0)
    swp = cuda_fwd_bwd(field, u, v, lat, lon, DT, go, interp, cfl=6.0)
    assert torch.equal(gen[0], swp[0])
    assert relmax(swp[1], gen[1]) < 2e-6, "grad_field"
    assert relmax(swp[2], gen[2]) < 1e-6 and relmax(swp[3], gen[3]) < 1e-6   # same formulas, other kernel
    again = cuda_fwd_bwd(field, u, v, lat, lon, DT, go, interp, cfl=6.0)
    assert torch.equal(again[1], swp[1]), "sweep is deterministic"
    ref = O.sl_advect_fwd_bwd(field, u, v, lat, lon, DT, go, interp)
    assert bad_fraction(swp[1], ref[1], 1e-4) < 1e-3


def test_sweep_contract_violation_falls_back():
    """cfl_cells is only a hint: planes whose displacement exceeds it are recomputed by the
    general path on the device, the result does not change."""
    H, W, B, V = 181, 360, 1, 6
    lat, lon, field, u, v, go = O.bench_inputs(H, W, B, V, True, DT, cells_sigma=1.0, cells_clip=2.0)
    u[:, 1] *= 4.0                      # planes 1 and 4 move up to 8 cells, the others at most 2
    v[:, 4] *= 4.0
    gen = cuda_fwd_bwd(field, u, v, lat, lon, DT, go, "bilinear", cfl=0.0)
    swp = cuda_fwd_bwd(field, u, v, lat, lon, DT, go, "bilinear", cfl=3.0)
    assert relmax(swp[1], gen[1]) < 2e-6
    assert torch.equal(swp[1][:, 1], gen[1][:, 1]) and torch.equal(swp[1][:, 4], gen[1][:, 4])   # fallback planes
    assert relmax(swp[2], gen[2]) < 1e-6 and relmax(swp[3], gen[3]) < 1e-6   # same formulas, other kernel


def test_sweep_adjoint_identity_full_size():
    H, W, B, V = 721, 1440, 1, 8
    lat, lon, field, u, v, go = O.bench_inputs(H, W, B, V, True, DT)
    pkg = P()
    geo = pkg.SLGeometry.from_grids(lat.cuda(), lon.cuda())
    f = field.cuda().requires_grad_(True)
    out = pkg.sl_advect(f, u.cuda(), v.cuda(), geo, DT, "bilinear", True, "fast", 6.0)
    out.backward(go.cuda())
    lhs = (out.detach().double() * go.cuda().double()).sum()
    rhs = (f.grad.double() * field.cuda().double()).sum()
    pkg.check_status()
    assert abs(float(lhs - rhs)) / abs(float(lhs)) < 1e-5


def test_lat_band_sweep_matches_global():
    """Bands + fused sweep (different task plan per band, hence a different fixed summation order)
    agree with the global call to rounding; grad_u / grad_v are unaffected by the decomposition."""
    from paradis_model_b200 import halo
    H, W, B, V = 181, 360, 1, 3
    lat, lon, field, u, v, go = O.bench_inputs(H, W, B, V, True, DT, cells_sigma=1.0, cells_clip=2.0)
    pkg = P()
    geo = pkg.SLGeometry.from_grids(lat.cuda(), lon.cuda())
    fc, uc, vc, gc = [t.cuda() for t in (field, u, v, go)]
    cfl = 3.0
    gfull = torch.ops.paradis.sl_advect_backward(gc, fc, uc, vc, geo.tables, geo.scalars, DT, 1, True, 0,
                                                 geo.windows, cfl, True, True)
    parts = []
    for rank in range(2):
        plan = halo.make_plan(H, W, rank, 2, cfl, "bilinear")
        own, ext = plan.windows()
        e = slice(ext[0], ext[0] + ext[1])
        gb = geo.band(own, ext, ext)
        parts.append(torch.ops.paradis.sl_advect_backward(
            gc[:, :, e].contiguous(), fc[:, :, e].contiguous(), uc[:, :, e].contiguous(), vc[:, :, e].contiguous(),
            gb.tables, gb.scalars, DT, 1, True, 0, gb.windows, cfl, True, True))
    pkg.check_status()
    for k in range(3):
        got = torch.cat([p[k] for p in parts], 2)
        assert relmax(got.cpu(), gfull[k].cpu()) < 2e-6


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_lat_band_nccl_two_gpus():
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29561", os.path.join(root, "bench.py"), "--gpus", "2", "--steps", "2",
           "--warmup", "1", "--decomp", "latband", "--workload", "c2", "--no-e2e", "--no-cpu"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    assert '"scaling": "strong"' in res.stdout


# ------------------------------------------------------------------ index mapping
@pytest.mark.parametrize("interp", ["bilinear", "bicubic"])
@pytest.mark.parametrize("poles,H,W", [(False, 128, 256), (True, 721, 1440)])
def test_exact_mode_coordinates_bit_identical_to_torch_cuda(poles, H, W, interp):
    """EXACT mode replays the reference's fp32 operation order: every intermediate of
    advection.py:82-96 and the sampler coordinates ATen floors are BIT-identical to the oracle's op
    replay executed by torch on the same GPU, hence so is the stencil index of every point."""
    from paradis_model_b200.ops import departure_coords
    lat, lon, field, u, v, go = O.bench_inputs(H, W, 1, 3, poles, DT)
    latc, lonc, uc, vc = lat.cuda(), lon.cuda(), u.cuda(), v.cuda()
    geo = P().SLGeometry.from_grids(latc, lonc)
    G = O.Geometry(latc, lonc)
    p = O.INTERP_PAD[interp]
    lat_d, lon_d = O.departure_latlon(uc, vc, G, DT)
    px, py = O.departure_pixels(uc, vc, G, DT)
    _, _, ix, iy = O.sampler_coords(px, py, H, W, p)
    got = departure_coords(uc, vc, geo, DT, interp, "exact")
    assert torch.equal(got[:, :, 9], lat_d) and torch.equal(got[:, :, 10], lon_d)
    assert torch.equal(got[:, :, 0], ix) and torch.equal(got[:, :, 1], iy)
    fast = departure_coords(uc, vc, geo, DT, interp, "fast")
    flips = ((fast[:, :, 0].floor() != ix.floor()) | (fast[:, :, 1].floor() != iy.floor())).float().mean().item()
    assert flips < 2e-4                       # fast math: ~1 ulp of ix, isolated cell-edge flips only
    assert (fast[:, :, 0] - ix).abs().max().item() < 1e-3 and (fast[:, :, 1] - iy).abs().max().item() < 1e-3


# ------------------------------------------------------------------ framework integration
def _small_problem(H=48, W=96, B=2, V=4):
    lat, lon = O.make_grids(H, W, True)
    field = O.smooth_field(lat, lon, B, V).float().cuda()
    u, v = [t.float().cuda() for t in O.smooth_velocity(lat, lon, B, V, 1.5, DT)]
    geo = P().SLGeometry.from_grids(lat.cuda(), lon.cuda())
    return geo, field, u, v


def test_torch_compile_fullgraph():
    """trainer.py:261-267 compiles the model with fullgraph=True: the custom op must trace (fake
    implementation + registered autograd) without a graph break."""
    pkg = P()
    geo, field, u, v = _small_problem()

    def fn(f, uu, vv):
        return pkg.sl_advect(f * 1.5, uu, vv, geo, DT, "bilinear").sin().sum()

    f1, u1, v1 = [t.clone().requires_grad_(True) for t in (field, u, v)]
    fn(f1, u1, v1).backward()
    f2, u2, v2 = [t.clone().requires_grad_(True) for t in (field, u, v)]
    cfn = torch.compile(fn, fullgraph=True, dynamic=False)
    cfn(f2, u2, v2).backward()
    assert relmax(f2.grad.cpu(), f1.grad.cpu()) < 1e-5
    assert relmax(u2.grad.cpu(), u1.grad.cpu()) < 1e-5 and relmax(v2.grad.cpu(), v1.grad.cpu()) < 1e-5


def test_bf16_autocast_and_checkpoint():
    """train.py:56 runs bf16-mixed; paradis.py:63-70 wraps layers in non-reentrant checkpoints."""
    from torch.utils.checkpoint import checkpoint
    pkg = P()
    geo, field, u, v = _small_problem()
    f32 = pkg.sl_advect(field, u, v, geo, DT, "bicubic")
    with torch.autocast("cuda", dtype=torch.bfloat16):
        fb, ub, vb = [t.bfloat16().requires_grad_(True) for t in (field, u, v)]
        out = checkpoint(lambda a, b, c: pkg.sl_advect(a, b, c, geo, DT, "bicubic"), fb, ub, vb, use_reentrant=False)
    assert out.dtype == torch.float32                       # as the reference's grid_sample under autocast
    assert relmax(out.cpu(), f32.cpu()) < 3e-2              # bf16 inputs
    out.sum().backward()
    assert fb.grad.dtype == torch.bfloat16 and ub.grad.dtype == torch.bfloat16
    assert torch.isfinite(fb.grad.float()).all() and torch.isfinite(ub.grad.float()).all()


# ------------------------------------------------------------------ ragged / degenerate shapes and error paths
@pytest.mark.parametrize("H,W,poles", [(5, 6, True), (7, 18, False), (33, 50, True), (16, 1442, False)])
@pytest.mark.parametrize("interp", ["bilinear", "bicubic"])
def test_ragged_shapes_scalar_paths(H, W, poles, interp):
    """W not a multiple of 4 (no float4 path), W not a multiple of 8 (byte-wise class scan), tiny H:
    forward and the general backward against the CPU oracle."""
    if interp == "bicubic" and H < 6:
        pytest.skip("padding 2 needs H >= 4 plus a row")
    B, V = 2, 3
    lat, lon = O.make_grids(H, W, poles)
    g = torch.Generator().manual_seed(H * 1000 + W)
    field = torch.randn(B, V, H, W, generator=g)
    sig = 1.2 * math.pi / H / DT
    u = (torch.randn(B, V, H, W, generator=g) * sig).clamp_(-2 * sig, 2 * sig)
    v = (torch.randn(B, V, H, W, generator=g) * sig).clamp_(-2 * sig, 2 * sig)
    go = torch.randn(B, V, H, W, generator=g)
    ref = O.sl_advect_fwd_bwd(field, u, v, lat, lon, DT, go, interp)
    got = cuda_fwd_bwd(field, u, v, lat, lon, DT, go, interp, "exact")
    # white noise against the CPU oracle: one ulp of ix at W = 1442 is 1.2e-4 cell, and EXACT follows
    # torch-CUDA's scalar-divide semantics rather than the CPU kernel's
    assert relmax(got[0], ref[0]) < 1e-3 and bad_fraction(got[0], ref[0], 1e-4) < 2e-2
    assert relmax(got[1], ref[1]) < 1e-3 and bad_fraction(got[1], ref[1], 1e-4) < 2e-2
    assert bad_fraction(got[2], ref[2], 1e-3) < 5e-3 and bad_fraction(got[3], ref[3], 1e-3) < 5e-3


def test_error_paths_raise():
    pkg = P()
    lat, lon = O.make_grids(8, 16, True)
    geo = pkg.SLGeometry.from_grids(lat.cuda(), lon.cuda())
    x = torch.zeros(1, 2, 8, 16, device="cuda")
    with pytest.raises(RuntimeError, match="shape mismatch"):
        pkg.sl_advect(x[:, :, :7], x, x, geo, DT)
    with pytest.raises(KeyError):
        pkg.sl_advect(x, x, x, geo, DT, "nearest")
    with pytest.raises(ValueError, match="separable"):
        pkg.SLGeometry.from_grids(lat.cuda() + lon.cuda(), lon.cuda())
    lat7, lon7 = O.make_grids(8, 16, True)
    with pytest.raises(RuntimeError, match="even"):
        g7 = pkg.SLGeometry(torch.zeros(2 * 8 + 15, device="cuda"), [0, 1, 0, 1], 8, 15)
        y = torch.zeros(1, 1, 8, 15, device="cuda")
        pkg.sl_advect(y, y, y, g7, DT)
    # non-contiguous field (channel slice of a wider tensor) is accepted
    wide = torch.randn(1, 4, 8, 16, device="cuda")
    a = pkg.sl_advect(wide[:, 1:3], x, x, geo, DT)
    b = pkg.sl_advect(wide[:, 1:3].contiguous(), x, x, geo, DT)
    assert torch.equal(a, b)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_fused_peer_halo_matches_nccl_and_single_gpu():
    """Latitude bands on 2 GPUs: the field halo read in place over NVLink peer memory (fused into the
    gather) is bit-identical to the NCCL-assembled halo and to the single-GPU result."""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29571", os.path.join(root, "tools", "check_p2p.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, (res.stdout[-1500:], res.stderr[-1500:])
    assert res.stdout.count("p2p==nccl True; fwd==single-GPU True") == 2


def test_cuda_graph_capture_replays_bit_identically():
    """Forward + fused backward (sweep on the caller's stream, polar caps on the library's side streams)
    captured in one CUDA graph: the replay reproduces the eager result bit for bit."""
    from paradis_model_b200.ops import RawAdvection
    pkg = P()
    H, W, B, V = 181, 360, 1, 3
    lat, lon, field, u, v, go = O.bench_inputs(H, W, B, V, True, DT)
    geo = pkg.SLGeometry.from_grids(lat.cuda(), lon.cuda())
    f, uu, vv, g = [t.cuda() for t in (field, u, v, go)]
    R = RawAdvection(geo, B, V, "bilinear", True, "fast", 6.0)

    def step():
        R.forward(f, uu, vv, DT)
        R.backward(g, f, uu, vv, DT, 3)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    ref = [t.clone() for t in (R.out, R.gfield, R.gu, R.gv)]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    for t in (R.out, R.gfield, R.gu, R.gv):
        t.zero_()
    graph.replay()
    torch.cuda.synchronize()
    for a, b in zip(ref, (R.out, R.gfield, R.gu, R.gv)):
        assert torch.equal(a, b)


# ------------------------------------------------------------------ randomised robustness of the backward paths
def test_random_shapes_sweep_vs_general_vs_oracle():
    """40 random configurations (mesh size, pole layout, interpolation, displacement scale, cfl hint, windows):
    fused sweep and general path must agree to rounding whatever the planner decides (bands, halos, caps,
    fallback planes), and both must match the CPU oracle."""
    import random
    rng = random.Random(1234)
    pkg = P()
    for trial in range(40):
        H = rng.choice([48, 61, 90, 128, 181, 240])
        W = rng.choice([160, 192, 256, 360, 512])
        poles = rng.random() < 0.5
        interp = rng.choice(["bilinear", "bilinear", "bicubic"])
        B, V = rng.choice([(1, 2), (2, 2), (1, 5)])
        clip = rng.choice([0.5, 1.5, 3.0, 6.0])
        cfl = rng.choice([1.0, 2.0, 4.0, 8.0, 12.0])          # sometimes too small: exercises the fallback
        lat, lon, field, u, v, go = O.bench_inputs(H, W, B, V, poles, DT, seed=trial, cells_sigma=clip / 2,
                                                   cells_clip=clip)
        gen = cuda_fwd_bwd(field, u, v, lat, lon, DT, go, interp, cfl=0.0)
        swp = cuda_fwd_bwd(field, u, v, lat, lon, DT, go, interp, cfl=cfl)
        tag = f"trial {trial}: {H}x{W} poles={poles} {interp} B{B} V{V} clip={clip} cfl={cfl}"
        assert torch.equal(gen[0], swp[0]), tag
        assert relmax(swp[1], gen[1]) < 5e-6, tag
        assert relmax(swp[2], gen[2]) < 1e-5 and relmax(swp[3], gen[3]) < 1e-5, tag
        if trial % 4 == 0:
            ref = O.sl_advect_fwd_bwd(field, u, v, lat, lon, DT, go, interp)
            assert bad_fraction(swp[0], ref[0], 2e-4) < 5e-3, tag
            assert bad_fraction(swp[1], ref[1], 2e-4) < 5e-3, tag


# ------------------------------------------------------------------ GeoCyclic padding fused into the depthwise conv
def test_dwconv_golden_from_reference_sepconv():
    import os
    from conftest import GOLDEN
    z = np.load(os.path.join(GOLDEN, "sepconv_depthwise.npz"))
    pkg = P()
    for k in (3, 5, 7):
        x = torch.from_numpy(z[f"k{k}_x"]).cuda().requires_grad_(True)
        w = torch.from_numpy(z[f"k{k}_w"]).cuda().requires_grad_(True)
        y = pkg.geocyclic_dwconv(x, w)
        y.backward(torch.from_numpy(z[f"k{k}_gy"]).cuda())
        assert relmax(y.detach().cpu(), torch.from_numpy(z[f"k{k}_y"])) < 1e-5, k
        assert relmax(x.grad.cpu(), torch.from_numpy(z[f"k{k}_gx"])) < 1e-5, k
        assert relmax(w.grad.cpu(), torch.from_numpy(z[f"k{k}_gw"])) < 1e-5, k


@pytest.mark.parametrize("k", [3, 5, 7])
@pytest.mark.parametrize("shape", [(2, 5, 33, 64), (1, 3, 128, 256), (1, 2, 721, 1440), (3, 4, 9, 38)])
def test_dwconv_vs_oracle(shape, k):
    """fwd / grad input / grad weight / grad bias against pad + F.conv2d (blocks.py:112-114) on the CPU;
    the backward is deterministic."""
    B, C, H, W = shape
    g = torch.Generator().manual_seed(k * 100 + H)
    x = torch.randn(B, C, H, W, generator=g)
    w = torch.randn(C, 1, k, k, generator=g) / k
    b = torch.randn(C, generator=g)
    gy = torch.randn(B, C, H, W, generator=g)
    xr, wr, br = [t.clone().requires_grad_(True) for t in (x, w, b)]
    O.geocyclic_depthwise(xr, wr, br).backward(gy)
    res = []
    for _ in range(2):
        xc, wc, bc = [t.cuda().requires_grad_(True) for t in (x, w, b)]
        y = P().geocyclic_dwconv(xc, wc, bc)
        y.backward(gy.cuda())
        res.append((y.detach().cpu(), xc.grad.cpu(), wc.grad.cpu(), bc.grad.cpu()))
    y_ref = O.geocyclic_depthwise(x, w, b)
    assert relmax(res[0][0], y_ref) < 1e-5
    assert relmax(res[0][1], xr.grad) < 1e-5
    assert relmax(res[0][2], wr.grad) < 1e-4 and relmax(res[0][3], br.grad) < 1e-4
    assert all(torch.equal(a, b2) for a, b2 in zip(res[0], res[1]))


def test_blocks_drop_ins_match_reference_golden():
    """paradis_model_b200.blocks.SepConv / PhysicalDownsample against outputs of the reference's classes."""
    import os
    from conftest import GOLDEN
    from paradis_model_b200 import blocks
    z = np.load(os.path.join(GOLDEN, "sepconv_depthwise.npz"))
    for stride in (1, 2, 4):
        x = torch.from_numpy(z[f"down_s{stride}_x"]).cuda()
        y = blocks.PhysicalDownsample(stride=stride).cuda()(x)
        ref = torch.from_numpy(z[f"down_s{stride}_y"])
        assert y.shape == ref.shape and relmax(y.cpu(), ref) < 1e-5, stride
    m = blocks.SepConv(3, 4, (12, 16), kernel_size=5).cuda()
    assert sorted(m.state_dict()) == ["depthwise.weight", "pointwise.bias", "pointwise.weight"]
    with torch.no_grad():
        m.depthwise.weight.copy_(torch.from_numpy(z["k5_w"]))
    x = torch.from_numpy(z["k5_x"]).cuda()
    dw = P().geocyclic_dwconv(x, m.depthwise.weight)
    assert relmax(dw.detach().cpu(), torch.from_numpy(z["k5_y"])) < 1e-5
    assert m(x).shape == (2, 4, 12, 16)


def test_faster_than_the_reference_ops_in_torch_cuda_eager():
    """SURVEY 8d's "more honest beat-this number": the reference's own op sequence (oracle op replay = what
    model/advection.py:129-169 executes) run by torch eager on the SAME GPU, at the benchmark's 0.25 degree / 64
    channel size, against this package's forward + backward.  The ratio is written to gpurun_out/ for DESIGN.md;
    the assertion only guards against a silent slow path."""
    import json, os
    H, W, B, V = 721, 1440, 1, 64
    lat, lon, field, u, v, go = O.bench_inputs(H, W, B, V, True, DT)
    latc, lonc, f, uu, vv, g = [t.cuda() for t in (lat, lon, field, u, v, go)]
    geo = P().SLGeometry.from_grids(latc, lonc)

    def ours():
        a, b, c = [t.detach().requires_grad_(True) for t in (f, uu, vv)]
        P().sl_advect(a, b, c, geo, DT, "bilinear", cfl_cells=6.0).backward(g)

    def ref():
        O.sl_advect_fwd_bwd(f, uu, vv, latc, lonc, DT, g, "bilinear")

    def timed(fn, n):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    t_ref, t_ours = timed(ref, 3), timed(ours, 10)
    mem = torch.cuda.max_memory_allocated() / 2**30
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/torch_cuda_eager.json", "w") as fh:
        json.dump({"workload": "c3 721x1440 V=64 B=1 bilinear fwd+bwd", "torch_cuda_eager_ms": t_ref,
                   "paradis_model_b200_autograd_op_ms": t_ours, "ratio": t_ref / t_ours,
                   "peak_mem_GiB_both": mem}, fh)
    assert t_ours * 3 < t_ref


@pytest.mark.parametrize("stride", [1, 2, 4])
def test_physical_downsample_strided_kernel(stride):
    """PhysicalDownsample (model/blocks.py:57-71) as one strided kernel: forward and input gradient against
    GeoCyclic pad 2 + avg_pool2d on the same GPU; only the strided outputs are computed."""
    import torch.nn.functional as F
    from paradis_model_b200 import blocks
    g = torch.Generator().manual_seed(stride)
    x = torch.randn(2, 3, 33, 64, generator=g).cuda()
    gy_shape = (2, 3, (33 - 1) // stride + 1, (64 - 1) // stride + 1)
    gy = torch.randn(gy_shape, generator=g).cuda()
    xr = x.clone().requires_grad_(True)
    yr = F.avg_pool2d(O.geocyclic_pad(xr, 2), 5, stride, count_include_pad=False)
    yr.backward(gy)
    xc = x.clone().requires_grad_(True)
    y = blocks.PhysicalDownsample(stride=stride).cuda()(xc)
    y.backward(gy)
    assert y.shape == yr.shape and relmax(y.detach().cpu(), yr.detach().cpu()) < 1e-6
    assert relmax(xc.grad.cpu(), xr.grad.cpu()) < 1e-5


def test_dwconv_rejects_meshes_where_both_caps_fold_onto_one_row():
    with pytest.raises(RuntimeError, match="mesh too small"):
        P().geocyclic_dwconv(torch.zeros(1, 2, 5, 16, device="cuda"), torch.zeros(2, 1, 5, 5, device="cuda"))


@pytest.mark.gpu
@pytest.mark.parametrize("W", [64, 30])
def test_halo_pack_matches_slices(W):
    """paradis_halo_pack: first / last h rows of several band tensors (one of them a strided channel view) into
    box[n][2][planes][h][W] in one launch -- compared with plain slicing."""
    from paradis_model_b200.ops import halo_pack
    B, V, rows, h = 2, 3, 17, 5
    g = torch.Generator().manual_seed(3)
    a = torch.randn(B, V, rows, W, generator=g).cuda()
    uv = torch.randn(B, 2 * V, rows, W, generator=g).cuda()
    u, v = uv[:, :V], uv[:, V:]                       # model/paradis.py:235-237: views of one tensor
    box = torch.full((2, 3, 2, B * V, h, W), -7.0, device="cuda")
    halo_pack([a, u, v], box[1], h)
    torch.cuda.synchronize()
    for k, t in enumerate((a, u, v)):
        assert torch.equal(box[1, k, 0], t[:, :, :h].reshape(B * V, h, W))
        assert torch.equal(box[1, k, 1], t[:, :, rows - h:].reshape(B * V, h, W))
    assert bool((box[0] == -7.0).all())
    with pytest.raises(RuntimeError):
        halo_pack([a.transpose(2, 3)], box[0], h)
