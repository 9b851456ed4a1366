import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import paradis_model_b200 as P
from paradis_model_b200 import synthetic as S
from paradis_model_b200.ops import RawAdvection
V = int(sys.argv[1]) if len(sys.argv) > 1 else 64
H, W, B = 721, 1440, 1
lat, lon = S.make_grids(H, W, True)
geo = P.SLGeometry.from_grids(lat.cuda(), lon.cuda())
f, u, v, g = [t.cuda() for t in S.white_noise_inputs(H, W, B, V)]
R = RawAdvection(geo, B, V, "bilinear", True, "fast", 6.0)
for _ in range(2):
    R.backward(g, f, u, v, S.DT_DEFAULT, 3)
torch.cuda.synchronize()
planes = B * V
al = lambda n: (n + 255) // 256 * 256
off = al(planes * 8) * 2 + al(planes * 4)
print("flags:", R.ws_b[off:off + planes].tolist())
print("reach:", R.ws_b[al(planes*8)*2: al(planes*8)*2 + planes*4].view(torch.int32).tolist()[:16])
P.check_status()
