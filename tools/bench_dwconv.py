"""Micro-benchmark (GPU box): fused GeoCyclic depthwise conv vs pad + torch depthwise conv (cuDNN / native)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import paradis_model_b200 as P

def ref_pad(x, p):   # restatement of model/padding.py:26-37 with torch ops (roll / flip / cat), runs on the GPU
    W = x.shape[3]
    top = torch.roll(x[:, :, 1:p + 1], W // 2, 3).flip(2)
    bot = torch.roll(x[:, :, -(p + 1):-1], W // 2, 3).flip(2)
    x = torch.cat([top, x, bot], 2)
    return torch.cat([x[..., -p:], x, x[..., :p]], 3)

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for (B, C, H, W) in [(1, 64, 721, 1440), (8, 256, 128, 256)]:
    for k in (3, 5, 7):
        x = torch.randn(B, C, H, W, device="cuda", requires_grad=True)
        w = (torch.randn(C, 1, k, k, device="cuda") / k).requires_grad_(True)
        gy = torch.randn(B, C, H, W, device="cuda")
        def ref_fwd(): return F.conv2d(ref_pad(x, (k - 1) // 2), w, None, groups=C)
        def our_fwd(): return P.geocyclic_dwconv(x, w)
        def ref_fb():
            x.grad = w.grad = None; ref_fwd().backward(gy)
        def our_fb():
            x.grad = w.grad = None; our_fwd().backward(gy)
        with torch.no_grad():
            tr, to = timeit(ref_fwd), timeit(our_fwd)
        trb, tob = timeit(ref_fb, 10), timeit(our_fb, 10)
        nbytes = 8 * x.numel()
        print(f"[{B},{C},{H},{W}] k={k}: fwd torch {tr:.3f} ms, fused {to:.3f} ms ({nbytes / to / 1e6:.0f} GB/s of 8 B/elem); "
              f"fwd+bwd torch {trb:.3f} ms, fused {tob:.3f} ms", flush=True)
