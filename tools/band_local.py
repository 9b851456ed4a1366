"""Single-GPU timing of one latitude band of the N-way split with all halo rows held LOCALLY (the NCCL-assembled layout):
separates the cost of the band structure (halo rows, thinner CTA segments) from the cost of peer-memory access.
usage: band_local.py [N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import paradis_model_b200 as P
from paradis_model_b200 import halo, synthetic as S
from paradis_model_b200.ops import RawAdvection
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
H, W, B, V, cfl = 721, 1440, 1, 64, 6.0
lat, lon = S.make_grids(H, W, True)
geo = P.SLGeometry.from_grids(lat.cuda(), lon.cuda())
full = S.white_noise_inputs(H, W, B, V)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for rank in range(N):
    plan = halo.make_plan(H, W, rank, N, cfl, "bilinear", balance=True)
    own, ext = plan.windows()
    e = slice(ext[0], ext[0] + ext[1]); o = slice(own[0], own[0] + own[1])
    f, u, v, g = [x[:, :, e].contiguous().cuda() for x in full]
    Rf = RawAdvection(geo.band(own, own, ext), B, V, "bilinear", True, "fast", cfl)
    Rb = RawAdvection(geo.band(own, ext, ext), B, V, "bilinear", True, "fast", cfl)
    uo, vo = [x[:, :, o].contiguous().cuda() for x in full[1:3]]
    tf = t(lambda: Rf.forward(f, uo, vo, S.DT_DEFAULT))
    tb = t(lambda: Rb.backward(g, f, u, v, S.DT_DEFAULT, 3))
    print(f"N={N} rank {rank}: own rows {own[1]} ext rows {ext[1]}: fwd {tf:.3f} ms bwd {tb:.3f} ms (ideal share of N=1: {own[1]/H*0.384:.3f} / {own[1]/H*1.811:.3f})", flush=True)
P.check_status()
