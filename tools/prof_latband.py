"""torch.profiler view of one latitude-band step (torchrun, rank 0 prints)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
import paradis_model_b200 as P
from paradis_model_b200 import halo, synthetic as S
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
H, W, B, V, cfl = 721, 1440, 1, 64, 6.0
lat, lon = S.make_grids(H, W, True)
geo = P.SLGeometry.from_grids(lat.to(dev), lon.to(dev))
plan = halo.make_plan(H, W, rank, world, cfl, "bilinear")
full = S.white_noise_inputs(H, W, B, V)
sl = slice(plan.row0, plan.row0 + plan.rows)
f, u, v, g = [t[:, :, sl].contiguous().to(dev) for t in full]
del full
peer = halo.PeerHalo(plan, B, V, dev)
def step():
    ff, uu, vv = f.requires_grad_(True), u.requires_grad_(True), v.requires_grad_(True)
    out = halo.lat_band_advect(ff, uu, vv, geo, plan, S.DT_DEFAULT, "bilinear", True, "fast", cfl, None, peer)
    out.backward(g)
    ff.grad = uu.grad = vv.grad = None
for _ in range(5): step()
torch.cuda.synchronize(); dist.barrier()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
if rank == 0:
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
dist.barrier(); dist.destroy_process_group()
