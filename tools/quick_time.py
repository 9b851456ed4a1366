"""Quick per-phase timing on the GPU box (not the bench contract; see bench.py)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import paradis_model_b200 as P
from paradis_model_b200 import synthetic as S
from paradis_model_b200.ops import RawAdvection

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

cases = [("C3", 721, 1440, 1, 64, True), ("C2", 128, 256, 8, 64, False)]
for name, H, W, B, V, poles in cases:
    lat, lon = S.make_grids(H, W, poles)
    geo = P.SLGeometry.from_grids(lat.cuda(), lon.cuda())
    f, u, v, g = [t.cuda() for t in S.white_noise_inputs(H, W, B, V)]
    pts = B * V * H * W
    for interp in ("bilinear", "bicubic"):
        for math in ("fast", "exact"):
            R = RawAdvection(geo, B, V, interp, True, math, 0.0)
            tf = timeit(lambda: R.forward(f, u, v, S.DT_DEFAULT))
            ta = timeit(lambda: R.backward(g, f, u, v, S.DT_DEFAULT, 1))
            tg = timeit(lambda: R.backward(g, f, u, v, S.DT_DEFAULT, 2))
            R2 = RawAdvection(geo, B, V, interp, True, math, 6.0)
            ts = timeit(lambda: R2.backward(g, f, u, v, S.DT_DEFAULT, 3))
            tot = tf + ts
            print(f"{name} {interp:8s} {math:5s}: fwd {tf:.3f} ms ({16*pts/tf/1e6:.0f} GB/s) arrival {ta:.3f} ms "
                  f"gather {tg:.3f} ms | fused bwd {ts:.3f} ms ({28*pts/ts/1e6:.0f} GB/s) total {tot:.3f} ms -> "
                  f"{pts/tot/1e6:.1f} Gpt/s, {44*pts/tot/1e6:.0f} GB/s", flush=True)
    P.check_status()
