import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as sm
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
buf = sm.empty(1024, dtype=torch.float32, device=f"cuda:{local}")
hdl = sm.rendezvous(buf, group=dist.group.WORLD)
buf.fill_(float(rank + 1))
hdl.barrier(channel=0)
print(rank, "ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal_pad_ptrs" in dir(hdl), type(hdl).__name__)
peer = hdl.get_buffer((rank + 1) % world, (1024,), torch.float32)
print(rank, "peer value", float(peer[0]))
hdl.barrier(channel=0)
dist.destroy_process_group()
