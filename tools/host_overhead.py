import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import paradis_model_b200 as P
from paradis_model_b200 import synthetic as S
from paradis_model_b200.ops import RawAdvection
H, W, B, V = 32, 64, 1, 4
lat, lon = S.make_grids(H, W, True)
geo = P.SLGeometry.from_grids(lat.cuda(), lon.cuda())
f, u, v, g = [t.cuda() for t in S.white_noise_inputs(H, W, B, V)]
def autograd_step():
    ff, uu, vv = f.requires_grad_(True), u.requires_grad_(True), v.requires_grad_(True)
    out = P.sl_advect(ff, uu, vv, geo, S.DT_DEFAULT, "bilinear")
    out.backward(g)
    ff.grad = uu.grad = vv.grad = None
R = RawAdvection(geo, B, V, "bilinear", True, "fast", 6.0)
def raw_step():
    R.forward(f, u, v, S.DT_DEFAULT); R.backward(g, f, u, v, S.DT_DEFAULT, 3)
for name, fn in (("autograd (torch.library op)", autograd_step), ("raw C-ABI (ctypes)", raw_step)):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200): fn()
    torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t0) / 200 * 1e6:.0f} us per fwd+bwd at {H}x{W} (host-bound)")
