"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the semi-Lagrangian advection path.

Two independent restatements of the reference algorithm live here:

* :func:`sl_advect` replays the reference's fp32 operation order
  (model/advection.py:129-169) with plain torch ops and hands the sampling to
  ``torch.nn.functional.grid_sample`` exactly like the reference call site
  (model/advection.py:161-167).  Gradients come from autograd.  On the same
  device it is bit-identical to the reference module (pinned by
  tests/test_oracle_vs_reference.py and by the fixtures in tests/golden/).
  This is also the "port" that bench.py times as the CPU baseline.

* :func:`sl_advect_explicit` is the closed form the CUDA kernels implement:
  explicit taps through the GeoCyclic index map, explicit Jacobian of the
  rotated-pole transform and an explicit adjoint scatter.  No autograd, no
  grid_sample.  It is checked against :func:`sl_advect` in fp64.

The arithmetic of ``grid_sample`` itself is a third-party dependency of the
reference (PyTorch, unpinned ``torch>=2.0.0`` in requirements.txt:2; de-facto
pin = torch 2.11.0 of this image).  Its published algorithm
(ATen/native/GridSampler.h:27-36, 205-297; ATen/native/UpSample.h:398-423) is
restated in :func:`_cubic_weights` / :func:`sl_advect_explicit`.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

INTERP_PAD = {"bilinear": 1, "bicubic": 2}


# --------------------------------------------------------------------------
# GeoCyclic padding (reference: model/padding.py:11-39)
# --------------------------------------------------------------------------
def geocyclic_source_index(H: int, W: int, p: int) -> tuple[np.ndarray, np.ndarray]:
    """Integer source map of the padded plane.

    Returns ``(src_row[Hp, Wp], src_col[Hp, Wp])`` such that
    ``padded[R, C] == x[src_row[R, C], src_col[R, C]]``.

    model/padding.py:26-31: the cap rows are rows ``1..p`` (resp.
    ``H-1-p..H-2``) rolled by ``W/2`` and flipped, i.e. a reflection about the
    first (last) row that *excludes* that row, shifted by 180 degrees.
    model/padding.py:35-37: longitude is periodic.
    """
    if W % 2:
        raise AssertionError("Number of longitude points must be even")
    Hp, Wp = H + 2 * p, W + 2 * p
    i = np.arange(Hp, dtype=np.int64)[:, None] - p
    j = np.arange(Wp, dtype=np.int64)[None, :] - p
    north = i < 0
    south = i >= H
    row = np.where(north, -i, np.where(south, 2 * (H - 1) - i, i))
    shift = np.where(north | south, W // 2, 0)
    col = np.mod(j - shift, W)
    return np.broadcast_to(row, (Hp, Wp)).copy(), col.astype(np.int64)


def geocyclic_pad(x: torch.Tensor, p: int) -> torch.Tensor:
    """Gather-based restatement of ``GeoCyclicPadding(p)(x)``."""
    if p == 0:
        return x
    assert x.dim() == 4, "Input must be 4-dimensional [batch, channels, lat, lon]"
    H, W = x.shape[-2:]
    row, col = geocyclic_source_index(H, W, p)
    row_t = torch.from_numpy(row).to(x.device)
    col_t = torch.from_numpy(col).to(x.device)
    return x[:, :, row_t, col_t]


def geocyclic_pad_adjoint(g: torch.Tensor, p: int) -> torch.Tensor:
    """Adjoint of :func:`geocyclic_pad`: fold the pads back onto their sources."""
    if p == 0:
        return g
    Hp, Wp = g.shape[-2:]
    H, W = Hp - 2 * p, Wp - 2 * p
    row, col = geocyclic_source_index(H, W, p)
    flat = torch.from_numpy(row * W + col).reshape(-1).to(g.device)
    out = torch.zeros(g.shape[:2] + (H * W,), dtype=g.dtype, device=g.device)
    out.index_add_(2, flat, g.reshape(g.shape[0], g.shape[1], -1))
    return out.reshape(g.shape[0], g.shape[1], H, W)


def geocyclic_depthwise(x: torch.Tensor, weight: torch.Tensor, bias=None) -> torch.Tensor:
    """First two lines of SepConv.forward (model/blocks.py:112-114): GeoCyclic pad by (k-1)//2, then the
    depthwise k x k convolution (groups = channels, no conv padding)."""
    k = weight.shape[-1]
    return F.conv2d(geocyclic_pad(x, (k - 1) // 2), weight, bias, groups=x.shape[1])


# --------------------------------------------------------------------------
# Pole continuity (reference: model/advection.py:100-114)
# --------------------------------------------------------------------------
def pole_mean(x: torch.Tensor) -> torch.Tensor:
    """Rows 0 and H-1 are replaced by their zonal mean (unconditionally)."""
    y = x.clone()
    y[:, :, 0, :] = x[:, :, 0:1, :].mean(dim=3, keepdim=True).squeeze(-1)
    y[:, :, -1, :] = x[:, :, -1:, :].mean(dim=3, keepdim=True).squeeze(-1)
    return y


# --------------------------------------------------------------------------
# Geometry (reference: model/advection.py:56-72)
# --------------------------------------------------------------------------
class Geometry:
    """0-dim tensors in the dtype of the grids, as the reference registers them."""

    def __init__(self, lat_grid: torch.Tensor, lon_grid: torch.Tensor):
        H, W = lat_grid.shape
        self.H, self.W = H, W
        self.lat = lat_grid.reshape(1, 1, H, W)
        self.lon = lon_grid.reshape(1, 1, H, W)
        self.Hf = torch.tensor(float(H), dtype=lat_grid.dtype, device=lat_grid.device)
        self.Wf = torch.tensor(float(W), dtype=lat_grid.dtype, device=lat_grid.device)
        self.min_lat = lat_grid.min()
        self.min_lon = lon_grid.min()
        self.d_lat = lat_grid.max() - self.min_lat
        self.d_lon = lon_grid.max() - self.min_lon


def make_grids(H: int, W: int, poles: bool, dtype=torch.float32):
    """Synthetic ERA5-style grids (data/era5_dataset.py:178-182: fp64 deg2rad, then cast).

    ``poles=True``  : lat = linspace(-90, 90, H)            (0.25 deg ERA5, H = 721)
    ``poles=False`` : lat = -90 + 180/H * (i + 1/2)         (WB2 pole-less grids)
    lon = 360/W * j in both cases.
    """
    if poles:
        lat = np.linspace(-90.0, 90.0, H)
    else:
        lat = -90.0 + 180.0 / H * (np.arange(H) + 0.5)
    lon = 360.0 / W * np.arange(W)
    lat_g, lon_g = np.meshgrid(np.deg2rad(lat), np.deg2rad(lon), indexing="ij")
    return (torch.from_numpy(lat_g).to(dtype), torch.from_numpy(lon_g).to(dtype))


# --------------------------------------------------------------------------
# Departure points, reference operation order
# --------------------------------------------------------------------------
def departure_latlon(u, v, geo: Geometry, dt: float):
    """model/advection.py:131-136 + 74-98 (rotated pole -> geographic)."""
    lon_r = -u * dt
    lat_r = -v * dt
    s_lat_r, c_lat_r = torch.sin(lat_r), torch.cos(lat_r)
    s_lon_r, c_lon_r = torch.sin(lon_r), torch.cos(lon_r)
    s_lat_p, c_lat_p = torch.sin(geo.lat), torch.cos(geo.lat)

    s = s_lat_r * c_lat_p + c_lat_r * c_lon_r * s_lat_p
    lat_dep = torch.arcsin(torch.clamp(s, -1 + 1e-7, 1 - 1e-7))
    num = c_lat_r * s_lon_r
    den = c_lat_r * c_lon_r * c_lat_p - s_lat_r * s_lat_p
    lon_dep = geo.lon + torch.atan2(num, den)
    lon_dep = torch.remainder(lon_dep + 2 * torch.pi, 2 * torch.pi)
    return lat_dep, lon_dep


def departure_pixels(u, v, geo: Geometry, dt: float):
    """model/advection.py:138-139: unpadded pixel coordinates of the departure point."""
    lat_dep, lon_dep = departure_latlon(u, v, geo, dt)
    pix_x = (lon_dep - geo.min_lon) / geo.d_lon * (geo.Wf - 1.0)
    pix_y = (lat_dep - geo.min_lat) / geo.d_lat * (geo.Hf - 1.0)
    return pix_x, pix_y


def sampler_coords(pix_x, pix_y, H: int, W: int, p: int):
    """model/advection.py:143-150 followed by ATen's un-normalisation
    (GridSampler.h:27-36, align_corners=True): the coordinates the sampler
    actually floors.  Same fp32 round trip as reference + ATen."""
    Hp, Wp = H + 2 * p, W + 2 * p
    gx = 2.0 * ((pix_x + p) / float(Wp - 1)) - 1.0
    gy = 2.0 * ((pix_y + p) / float(Hp - 1)) - 1.0
    ix = ((gx + 1) / 2) * (Wp - 1)
    iy = ((gy + 1) / 2) * (Hp - 1)
    return gx, gy, ix, iy


def sl_advect(field, u, v, lat_grid, lon_grid, dt: float, interpolation: str = "bilinear",
              pole_fix: bool = True):
    """Operator core: reference model/advection.py:129-169 (projections excluded).

    field, u, v: [B, V, H, W]; returns [B, V, H, W].  Differentiable (autograd).
    """
    B, V, H, W = field.shape
    p = INTERP_PAD[interpolation]
    geo = Geometry(lat_grid, lon_grid)
    src = pole_mean(field) if pole_fix else field
    pix_x, pix_y = departure_pixels(u, v, geo, dt)
    padded = geocyclic_pad(src, p)
    gx, gy, _, _ = sampler_coords(pix_x, pix_y, H, W, p)
    grid = torch.stack([gx.reshape(B * V, H, W), gy.reshape(B * V, H, W)], dim=-1)
    out = F.grid_sample(padded.reshape(B * V, 1, H + 2 * p, W + 2 * p), grid,
                        align_corners=True, mode=interpolation, padding_mode="zeros")
    out = out.reshape(B, V, H, W)
    return pole_mean(out) if pole_fix else out


def sl_advect_fwd_bwd(field, u, v, lat_grid, lon_grid, dt, grad_out, interpolation="bilinear",
                      pole_fix=True):
    """One forward + backward through autograd.  Returns (out, gfield, gu, gv)."""
    f = field.detach().clone().requires_grad_(True)
    uu = u.detach().clone().requires_grad_(True)
    vv = v.detach().clone().requires_grad_(True)
    out = sl_advect(f, uu, vv, lat_grid, lon_grid, dt, interpolation, pole_fix)
    out.backward(grad_out)
    return out.detach(), f.grad, uu.grad, vv.grad


# --------------------------------------------------------------------------
# Explicit closed form (what the CUDA kernels compute)
# --------------------------------------------------------------------------
_A = -0.75


def _cubic_weights(t):
    """ATen/native/UpSample.h:398-423 (A=-0.75), taps at floor-1 .. floor+2."""
    def cc1(x):
        return ((_A + 2) * x - (_A + 3)) * x * x + 1

    def cc2(x):
        return ((_A * x - 5 * _A) * x + 8 * _A) * x - 4 * _A

    return [cc2(t + 1.0), cc1(t), cc1(1.0 - t), cc2((1.0 - t) + 1.0)]


def _cubic_weights_grad(t):
    """d(weight)/dt.  ATen/native/GridSampler.h:280-297 tabulates the NEGATIVE of
    these (its caller subtracts: ``gix -= value * coeff_grad * ...``)."""
    x0, x1, x2, x3 = -1 - t, -t, 1 - t, 2 - t
    return [-((-3 * _A * x0 - 10 * _A) * x0 - 8 * _A),
            -((-3 * (_A + 2) * x1 - 2 * (_A + 3)) * x1),
            -((3 * (_A + 2) * x2 - 2 * (_A + 3)) * x2),
            -((3 * _A * x3 - 10 * _A) * x3 + 8 * _A)]


def _stencil(ix, iy, interpolation):
    """Per-axis tap offsets, weights and weight derivatives."""
    x0, y0 = torch.floor(ix), torch.floor(iy)
    tx, ty = ix - x0, iy - y0
    if interpolation == "bilinear":
        offs = [0, 1]
        wx, wy = [1 - tx, tx], [1 - ty, ty]
        one = torch.ones_like(tx)
        dwx, dwy = [-one, one], [-one, one]
    else:
        offs = [-1, 0, 1, 2]
        wx, wy = _cubic_weights(tx), _cubic_weights(ty)
        dwx, dwy = _cubic_weights_grad(tx), _cubic_weights_grad(ty)
    return x0.long(), y0.long(), offs, wx, wy, dwx, dwy


def sl_advect_explicit(field, u, v, lat_grid, lon_grid, dt, grad_out=None,
                       interpolation="bilinear", pole_fix=True):
    """Closed-form forward and (if ``grad_out`` is given) backward.

    Forward : out = sum_taps w * field[src(R, C)]                      
    Backward: grad_u, grad_v through the Jacobian of the rotated-pole map,
              grad_field by the adjoint scatter through the GeoCyclic map.
    Works in the dtype of the inputs (use fp64 for formula checks).
    """
    B, V, H, W = field.shape
    p = INTERP_PAD[interpolation]
    Hp, Wp = H + 2 * p, W + 2 * p
    geo = Geometry(lat_grid, lon_grid)
    src = pole_mean(field) if pole_fix else field

    lon_r, lat_r = -u * dt, -v * dt
    sa, ca = torch.sin(lat_r), torch.cos(lat_r)          # phi'
    sb, cb = torch.sin(lon_r), torch.cos(lon_r)          # lambda'
    sp, cp = torch.sin(geo.lat), torch.cos(geo.lat)      # arrival point
    s = sa * cp + ca * cb * sp
    lo, hi = -1 + 1e-7, 1 - 1e-7
    s_c = torch.clamp(s, lo, hi)
    num = ca * sb
    den = ca * cb * cp - sa * sp
    lat_dep = torch.arcsin(s_c)
    lon_dep = torch.remainder(geo.lon + torch.atan2(num, den) + 2 * math.pi, 2 * math.pi)
    Ax = (geo.Wf - 1.0) / geo.d_lon
    Ay = (geo.Hf - 1.0) / geo.d_lat
    ix = (lon_dep - geo.min_lon) * Ax + p
    iy = (lat_dep - geo.min_lat) * Ay + p

    x0, y0, offs, wx, wy, dwx, dwy = _stencil(ix, iy, interpolation)
    row_map, col_map = geocyclic_source_index(H, W, p)
    flat_map = torch.from_numpy(row_map * W + col_map).to(field.device)      # [Hp, Wp]
    src_flat = src.reshape(B, V, H * W)

    out = torch.zeros_like(field)
    dsum_x = torch.zeros_like(field)
    dsum_y = torch.zeros_like(field)
    taps = []
    for a, oy in enumerate(offs):
        for b, ox in enumerate(offs):
            R, C = y0 + oy, x0 + ox
            ok = (R >= 0) & (R < Hp) & (C >= 0) & (C < Wp)          # padding_mode="zeros"
            idx = flat_map[R.clamp(0, Hp - 1), C.clamp(0, Wp - 1)]   # [B,V,H,W]
            val = torch.gather(src_flat, 2, idx.reshape(B, V, -1)).reshape(B, V, H, W)
            val = torch.where(ok, val, torch.zeros_like(val))
            out = out + wy[a] * wx[b] * val
            dsum_x = dsum_x + wy[a] * dwx[b] * val
            dsum_y = dsum_y + dwy[a] * wx[b] * val
            taps.append((idx, ok, wy[a] * wx[b]))
    res = pole_mean(out) if pole_fix else out
    if grad_out is None:
        return res

    def pole_mean_adjoint(g):
        gg = g.clone()
        gg[:, :, 0, :] = g[:, :, 0, :].mean(dim=-1, keepdim=True)
        gg[:, :, -1, :] = g[:, :, -1, :].mean(dim=-1, keepdim=True)
        return gg

    g = pole_mean_adjoint(grad_out) if pole_fix else grad_out
    gix, giy = g * dsum_x, g * dsum_y

    r2 = num * num + den * den
    dlam_db = (den * ca * cb + num * ca * sb * cp) / r2
    dlam_da = (-den * sa * sb + num * (sa * cb * cp + ca * sp)) / r2
    inside = ((s >= lo) & (s <= hi)).to(field.dtype)
    dphi_ds = inside / torch.sqrt(1 - s_c * s_c)
    ds_db = -ca * sb * sp
    ds_da = ca * cp - sa * cb * sp
    grad_u = -dt * (gix * Ax * dlam_db + giy * Ay * dphi_ds * ds_db)
    grad_v = -dt * (gix * Ax * dlam_da + giy * Ay * dphi_ds * ds_da)

    gsrc = torch.zeros(B, V, H * W, dtype=field.dtype, device=field.device)
    for idx, ok, w in taps:
        contrib = torch.where(ok, w * g, torch.zeros_like(g))
        gsrc.scatter_add_(2, idx.reshape(B, V, -1), contrib.reshape(B, V, -1))
    gsrc = gsrc.reshape(B, V, H, W)
    grad_field = pole_mean_adjoint(gsrc) if pole_fix else gsrc
    return res, grad_field, grad_u, grad_v


# --------------------------------------------------------------------------
# Synthetic inputs shared by tests and bench
# --------------------------------------------------------------------------
def smooth_field(lat_grid, lon_grid, B: int, V: int, seed: int = 0):
    """Band-limited analytic fields, different phase/amplitude per plane."""
    g = torch.Generator().manual_seed(seed)
    amp = 0.5 + torch.rand(B, V, 1, 1, generator=g, dtype=torch.float64)
    ph = 2 * math.pi * torch.rand(B, V, 1, 1, generator=g, dtype=torch.float64)
    la, lo = lat_grid.double()[None, None], lon_grid.double()[None, None]
    f = amp * (torch.sin(3 * lo + ph) * torch.cos(la) ** 3
               + torch.cos(5 * lo + 1 + ph) * torch.cos(la) ** 5 * torch.sin(2 * la))
    return f


def smooth_velocity(lat_grid, lon_grid, B: int, V: int, cells: float, dt: float, seed: int = 1):
    """Smooth (u, v) whose displacement is about ``cells`` latitude cells."""
    H = lat_grid.shape[0]
    g = torch.Generator().manual_seed(seed)
    ph = 2 * math.pi * torch.rand(2, B, V, 1, 1, generator=g, dtype=torch.float64)
    la, lo = lat_grid.double()[None, None], lon_grid.double()[None, None]
    scale = cells * (math.pi / H) / dt
    u = scale * (torch.cos(2 * lo + ph[0]) * torch.cos(la) + 0.3 * torch.sin(la + ph[1]))
    v = scale * (torch.sin(3 * lo + ph[1]) * torch.cos(2 * la) + 0.3 * torch.cos(lo + ph[0]))
    return u, v


def bench_inputs(H: int, W: int, B: int, V: int, poles: bool, dt: float, seed: int = 0,
                 cells_sigma: float = 2.0, cells_clip: float = 4.0):
    """SURVEY 8d synthetic inputs: field~N(0,1); u,v~N(0,sigma^2) with
    sigma = cells_sigma*dphi/dt clipped at +-cells_clip cells; grad_out~N(0,1)."""
    lat_grid, lon_grid = make_grids(H, W, poles)
    g = torch.Generator().manual_seed(seed)
    dphi = math.pi / H
    sigma, clip = cells_sigma * dphi / dt, cells_clip * dphi / dt
    field = torch.randn(B, V, H, W, generator=g)
    u = (torch.randn(B, V, H, W, generator=g) * sigma).clamp_(-clip, clip)
    v = (torch.randn(B, V, H, W, generator=g) * sigma).clamp_(-clip, clip)
    grad_out = torch.randn(B, V, H, W, generator=g)
    return lat_grid, lon_grid, field, u, v, grad_out
