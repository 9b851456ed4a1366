"""BASELINE config 4 under torchrun: one training step of the PARADIS assembly with the drop-in modules under
DistributedDataParallel (train.py:49), batch-sharded, bf16 autocast, at 1.40625 degrees.  Prints "ddp step ok"."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP

from oracle import paradis_assembly as A
from oracle.sl_oracle import make_grids
import paradis_model_b200 as P

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
H, W = 128, 256
cfg = A.default_cfg(interp="bicubic")
lat, lon = make_grids(H, W, False)
torch.manual_seed(0)
model = A.Assembly(A.FakeDataModule, cfg, lat, lon, dropin=True).to(dev)
ddp = DDP(model, device_ids=[local])
opt = torch.optim.AdamW(ddp.parameters(), lr=1e-3)
g = torch.Generator().manual_seed(100 + rank)          # every rank its own batch shard
x = torch.randn(1, 22, H, W, generator=g).to(dev)
tgt = torch.randn(1, 7, H, W, generator=g).to(dev)
with torch.autocast("cuda", dtype=torch.bfloat16):
    loss = torch.nn.functional.mse_loss(ddp(x).float(), tgt)
loss.backward()
opt.step()
P.check_status(dev)
# the all-reduced gradients keep the replicas identical
flat = torch.cat([p.detach().flatten() for p in model.parameters()])
ref = flat.clone()
dist.broadcast(ref, 0)
assert torch.equal(flat, ref), "replicas diverged"
if rank == 0:
    print(f"ddp step ok: loss {float(loss):.4f}, {flat.numel()} parameters identical on {dist.get_world_size()} ranks", flush=True)
dist.barrier()
dist.destroy_process_group()
