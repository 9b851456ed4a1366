"""Diagnostic (GPU box): where do grad_u / grad_v of the fused op differ from the oracle run by torch on the same GPU at
721x1440 with smooth fields?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import paradis_model_b200 as P
from paradis_model_b200.ops import departure_coords
from oracle import sl_oracle as O
DT = 21600 * 7.29212e-5 / 8
H, W, B, V = 721, 1440, 1, 3
lat, lon = O.make_grids(H, W, True)
field = O.smooth_field(lat, lon, B, V).float()
u, v = [t.float() for t in O.smooth_velocity(lat, lon, B, V, 4.0, DT)]
go = O.smooth_field(lat, lon, B, V, seed=7).float()
dev = [t.cuda() for t in (field, u, v, lat, lon, go)]
for interp in ("bilinear", "bicubic"):
    ref = O.sl_advect_fwd_bwd(dev[0], dev[1], dev[2], dev[3], dev[4], DT, dev[5], interp)
    geo = P.SLGeometry.from_grids(dev[3], dev[4])
    for math in ("fast", "exact"):
        f, uu, vv = [t.clone().requires_grad_(True) for t in dev[:3]]
        out = P.sl_advect(f, uu, vv, geo, DT, interp, True, math, 0.0)
        out.backward(dev[5])
        for name, a, b in (("gf", f.grad, ref[1]), ("gu", uu.grad, ref[2]), ("gv", vv.grad, ref[3])):
            d = (a - b).abs()
            rowmax = d.amax(dim=(0, 1, 3)) / b.abs().max()
            worst = torch.topk(rowmax, 6)
            i = int(d.argmax())
            pl, r, c = i // (H * W), (i // W) % H, i % W
            print(f"{interp} {math} {name}: max|ref| {float(b.abs().max()):.3e} worst rows {[(int(k), f'{float(x):.1e}') for x, k in zip(worst.values, worst.indices)]}"
                  f" argmax plane {pl} row {r} col {c}: ours {float(a.flatten()[i]):.6e} ref {float(b.flatten()[i]):.6e}")
        if math == "fast" and interp == "bilinear":
            i = int((uu.grad - ref[2]).abs().argmax())
            pl, r, c = i // (H * W), (i // W) % H, i % W
            dc = departure_coords(dev[1], dev[2], geo, DT, interp, "fast")[0, pl, :, r, c]
            de = departure_coords(dev[1], dev[2], geo, DT, interp, "exact")[0, pl, :, r, c]
            print(" coords fast :", [f"{float(x):.9g}" for x in dc])
            print(" coords exact:", [f"{float(x):.9g}" for x in de])
            print(" ref gu neighbourhood:", [f"{float(x):.4e}" for x in ref[2][0, pl, r, max(c-2,0):c+3]])
            print(" our gu neighbourhood:", [f"{float(x):.4e}" for x in uu.grad[0, pl, r, max(c-2,0):c+3]])
