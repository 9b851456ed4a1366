"""Diagnostic (GPU box): where do CUDA / fp32 oracle / fp64 oracle differ on smooth fields?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import sl_oracle as O
import paradis_model_b200 as P
DT = 21600 * 7.29212e-5 / 8
H, W, B, V = 181, 360, 2, 4
for interp in ("bilinear",):
    lat, lon = O.make_grids(H, W, True)
    lat64, lon64 = O.make_grids(H, W, True, torch.float64)
    field = O.smooth_field(lat, lon, B, V).float()
    u, v = [t.float() for t in O.smooth_velocity(lat, lon, B, V, 2.5, DT)]
    go = O.smooth_field(lat, lon, B, V, seed=7).float()
    r32 = O.sl_advect_fwd_bwd(field, u, v, lat, lon, DT, go, interp)
    r64 = O.sl_advect_fwd_bwd(field.double(), u.double(), v.double(), lat64, lon64, DT, go.double(), interp)
    geo = P.SLGeometry.from_grids(lat.cuda(), lon.cuda())
    res = {}
    for mode in ("fast", "exact"):
        f, uu, vv = [t.cuda().requires_grad_(True) for t in (field, u, v)]
        out = P.sl_advect(f, uu, vv, geo, DT, interp, True, mode)
        out.backward(go.cuda())
        res[mode] = (out.detach().cpu(), f.grad.cpu(), uu.grad.cpu(), vv.grad.cpu())
    d = [t.cuda() for t in (field, u, v, lat, lon, go)]
    rg = [t.cpu() for t in O.sl_advect_fwd_bwd(d[0], d[1], d[2], d[3], d[4], DT, d[5], interp)]
    names = ["out", "gfield", "gu", "gv"]
    for k in range(4):
        sc = r64[k].abs().max()
        def rowerr(a):
            return ((a.double() - r64[k]).abs().amax(dim=(0, 1, 3)) / sc)
        e32, ef, ee, eg = rowerr(r32[k]), rowerr(res["fast"][k]), rowerr(res["exact"][k]), rowerr(rg[k])
        print(f"{interp} {names[k]}: vs fp64 max: cpu32 {e32.max():.2e} cuda-torch32 {eg.max():.2e} fast {ef.max():.2e} exact {ee.max():.2e}")
        worst = torch.argsort(ef, descending=True)[:6]
        print("   worst rows(fast):", [(int(r), f"{ef[r]:.1e}", f"cpu32 {e32[r]:.1e}", f"exact {ee[r]:.1e}") for r in worst])
        exd = (res["exact"][k] - rg[k]).abs().max() / sc
        print(f"   exact vs same-device torch: {exd:.2e}")
