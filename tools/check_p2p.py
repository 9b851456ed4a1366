"""2+ GPU check (torchrun): lat-band forward/backward with the fused peer-memory field halo must equal
the NCCL-assembled halo bit for bit, and both must match the single-GPU result on rank 0's band."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import paradis_model_b200 as P
from paradis_model_b200 import halo, synthetic as S
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
H, W, B, V, cfl = 181, 360, 1, 4, 3.0
lat, lon = S.make_grids(H, W, True)
geo = P.SLGeometry.from_grids(lat.to(dev), lon.to(dev))
full = S.white_noise_inputs(H, W, B, V, cells_sigma=1.0, cells_clip=2.0)
plan = halo.make_plan(H, W, rank, world, cfl, "bilinear")
sl = slice(plan.row0, plan.row0 + plan.rows)
res = {}
peer = halo.PeerHalo(plan, B, V, dev)
for name, pr in (("nccl", None), ("p2p", peer)):
    f, u, v, g = [t[:, :, sl].contiguous().to(dev) for t in full]
    f.requires_grad_(True); u.requires_grad_(True); v.requires_grad_(True)
    out = halo.lat_band_advect(f, u, v, geo, plan, S.DT_DEFAULT, "bilinear", True, "fast", cfl, None, pr)
    out.backward(g)
    P.check_status(dev)
    res[name] = (out.detach(), f.grad, u.grad, v.grad)
same = all(torch.equal(a, b) for a, b in zip(res["nccl"], res["p2p"]))
# single-GPU reference of this band
ff, uu, vv, gg = [t.to(dev).requires_grad_(True) for t in full[:3]] + [full[3].to(dev)]
o = P.sl_advect(ff, uu, vv, geo, S.DT_DEFAULT, "bilinear", True, "fast", cfl)
o.backward(gg)
ok_fwd = torch.equal(o.detach()[:, :, sl], res["p2p"][0])
err_gf = float((ff.grad[:, :, sl] - res["p2p"][1]).abs().max() / ff.grad.abs().max())
ok_gu = float((uu.grad[:, :, sl] - res["p2p"][2]).abs().max() / uu.grad.abs().max())
print(f"rank {rank}: p2p==nccl {same}; fwd==single-GPU {ok_fwd}; grad_field rel diff {err_gf:.2e}; grad_u rel diff {ok_gu:.2e}", flush=True)
assert same and ok_fwd and err_gf < 2e-6 and ok_gu < 1e-6
dist.barrier(); dist.destroy_process_group()
