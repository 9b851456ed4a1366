import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import paradis_model_b200 as P
from paradis_model_b200 import synthetic as S
H, W, B, V = 721, 1440, 1, 64
lat, lon = S.make_grids(H, W, True)
geo = P.SLGeometry.from_grids(lat.cuda(), lon.cuda())
h_in = S.white_noise_inputs(H, W, B, V, pin=True)
h_out = [torch.empty(B, V, H, W, pin_memory=True) for _ in range(4)]
for chunk in (2, 4, 8, 16, 32):
    scratch = P.host_fwd_bwd(geo, *h_in, *h_out, S.DT_DEFAULT, "bilinear", True, "fast", chunk, None, 6.0)
    t0 = time.perf_counter()
    for _ in range(4):
        scratch = P.host_fwd_bwd(geo, *h_in, *h_out, S.DT_DEFAULT, "bilinear", True, "fast", chunk, scratch, 6.0)
    dt = (time.perf_counter() - t0) / 4
    print(f"chunk_planes {chunk:3d}: {dt * 1e3:.2f} ms/step, {2 * 4 * h_in[0].numel() * 4 / dt / 1e9:.1f} GB/s PCIe (both directions)")
    del scratch
