import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import paradis_model_b200 as P
from paradis_model_b200 import synthetic as S, _lib
from paradis_model_b200.ops import RawAdvection
H, W, B, V = 721, 1440, 1, 64
lat, lon = S.make_grids(H, W, True)
geo = P.SLGeometry.from_grids(lat.cuda(), lon.cuda())
f, u, v, g = [t.cuda() for t in S.white_noise_inputs(H, W, B, V)]
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for interp in ("bilinear", "bicubic"):
    R0 = RawAdvection(geo, B, V, interp, True, "fast", 0.0)
    R = RawAdvection(geo, B, V, interp, True, "fast", 6.0)
    print(os.environ.get("PARADIS_SL_LIB", "default"), interp, "fwd %.3f arrival %.3f gather %.3f fused-bwd %.3f" % (
        t(lambda: R.forward(f, u, v, S.DT_DEFAULT)), t(lambda: R0.backward(g, f, u, v, S.DT_DEFAULT, 1)),
        t(lambda: R0.backward(g, f, u, v, S.DT_DEFAULT, 2)), t(lambda: R.backward(g, f, u, v, S.DT_DEFAULT, 3))))
