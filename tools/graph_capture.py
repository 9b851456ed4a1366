"""CUDA-graph capture of one forward + fused backward through the C ABI (side streams included)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import paradis_model_b200 as P
from paradis_model_b200 import synthetic as S
from paradis_model_b200.ops import RawAdvection
for (H, W, B, V, poles) in [(181, 360, 1, 4, True), (128, 256, 8, 64, False)]:
    lat, lon = S.make_grids(H, W, poles)
    geo = P.SLGeometry.from_grids(lat.cuda(), lon.cuda())
    f, u, v, g = [t.cuda() for t in S.white_noise_inputs(H, W, B, V)]
    R = RawAdvection(geo, B, V, "bilinear", True, "fast", 6.0)
    def step():
        R.forward(f, u, v, S.DT_DEFAULT); R.backward(g, f, u, v, S.DT_DEFAULT, 3)
    for _ in range(3): step()
    torch.cuda.synchronize()
    ref = [t.clone() for t in (R.out, R.gfield, R.gu, R.gv)]
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step()
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    for t in (R.out, R.gfield, R.gu, R.gv): t.zero_()
    graph.replay(); torch.cuda.synchronize()
    same = all(torch.equal(a, b) for a, b in zip(ref, (R.out, R.gfield, R.gu, R.gv)))
    def timeit(fn, n=200):
        for _ in range(10): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(n): fn()
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
    print(f"{H}x{W} B{B} V{V}: graph replay bit-identical {same}; eager {timeit(step):.0f} us/step, graph {timeit(graph.replay):.0f} us/step")
