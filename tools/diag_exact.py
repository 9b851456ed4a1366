"""Where does EXACT mode leave the reference's fp32 values?  Compare every intermediate of the
departure chain with torch-CUDA eager (oracle op replay) on the same GPU."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import sl_oracle as O
import paradis_model_b200 as P
from paradis_model_b200.ops import departure_coords
DT = 21600 * 7.29212e-5 / 8
for (H, W, poles) in [(128, 256, False), (721, 1440, True)]:
    lat, lon, field, u, v, go = O.bench_inputs(H, W, 1, 4, poles, DT)
    latc, lonc, uc, vc = lat.cuda(), lon.cuda(), u.cuda(), v.cuda()
    geo = P.SLGeometry.from_grids(latc, lonc)
    G = O.Geometry(latc, lonc)
    lon_r, lat_r = -uc * DT, -vc * DT
    sa, ca, sb, cb = torch.sin(lat_r), torch.cos(lat_r), torch.sin(lon_r), torch.cos(lon_r)
    sp, cp = torch.sin(G.lat), torch.cos(G.lat)
    s = sa * cp + ca * cb * sp
    num = ca * sb
    den = ca * cb * cp - sa * sp
    px, py = O.departure_pixels(uc, vc, G, DT)
    _, _, ix, iy = O.sampler_coords(px, py, H, W, 1)
    lat_d, lon_d = O.departure_latlon(uc, vc, G, DT)
    ref = [ix, iy, sa, ca, sb, cb, s, num, den, lat_d, lon_d]
    names = ["ix", "iy", "sin_lat'", "cos_lat'", "sin_lon'", "cos_lon'", "s", "num", "den", "lat_dep", "lon_dep"]
    for mode in ("exact", "fast"):
        got = departure_coords(uc, vc, geo, DT, "bilinear", mode)
        print(f"--- {H}x{W} {mode}")
        for k, n in enumerate(names):
            a, b = got[:, :, k], ref[k]
            neq = (a != b).float().mean().item()
            print(f"  {n:9s} mismatching {neq:9.2e}  max|diff| {(a - b).abs().max().item():.3e}")
        fl = ((got[:, :, 0].floor() != ix.floor()) | (got[:, :, 1].floor() != iy.floor())).float().mean().item()
        print(f"  floor(ix,iy) differs at {fl:.2e} of points")
    # pieces: lat = asin(clamp(s)), atan2
    sc = torch.clamp(s, -1 + 1e-7, 1 - 1e-7)
    # bisect the pixel chain with torch ops fed by OUR lat/lon (exact mode)
    got = departure_coords(uc, vc, geo, DT, "bilinear", "exact")
    la, lo = got[:, :, 9], got[:, :, 10]
    px2 = (lo - G.min_lon) / G.d_lon * (G.Wf - 1.0)
    py2 = (la - G.min_lat) / G.d_lat * (G.Hf - 1.0)
    _, _, ix2, iy2 = O.sampler_coords(px2, py2, H, W, 1)
    print("  pixel chain from our lat/lon via torch: ix mismatch", (ix2 != got[:, :, 0]).float().mean().item(),
          "iy mismatch", (iy2 != got[:, :, 1]).float().mean().item())
    at = torch.atan2(num, den)
    l1 = G.lon + at
    l2 = l1 + 2 * torch.pi
    l3 = torch.remainder(l2, 2 * torch.pi)
    print("  torch scalars: min_lon", G.min_lon.item(), "d_lon", G.d_lon.item(), "min_lat", G.min_lat.item(), "d_lat", G.d_lat.item())
    print("  geo scalars  :", geo.scalars)
