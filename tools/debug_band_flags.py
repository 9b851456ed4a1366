import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import paradis_model_b200 as P
from paradis_model_b200 import synthetic as S, halo
from paradis_model_b200.ops import RawAdvection
H, W, B, V, cfl = 721, 1440, 1, 8, 6.0
lat, lon = S.make_grids(H, W, True)
geo = P.SLGeometry.from_grids(lat.cuda(), lon.cuda())
full = [t.cuda() for t in S.white_noise_inputs(H, W, B, V)]
for rank in range(2):
    plan = halo.make_plan(H, W, rank, 2, cfl, "bilinear")
    own, ext = plan.windows()
    e = slice(ext[0], ext[0] + ext[1])
    gb = geo.band(own, ext, ext)
    f, u, v, g = [t[:, :, e].contiguous() for t in full]
    R = RawAdvection(gb, B, V, "bilinear", True, "fast", cfl)
    R.backward(g, f, u, v, S.DT_DEFAULT, 3)
    torch.cuda.synchronize()
    planes = B * V
    al = lambda n: (n + 255) // 256 * 256
    off = al(planes * 8) * 2 + 3 * al(planes * 4)
    print("rank", rank, "own", own, "ext", ext, "flags:", R.ws_b[off:off + planes].tolist())
P.check_status()
