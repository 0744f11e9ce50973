import os, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import torch.distributed._symmetric_memory as symm_mem
t = symm_mem.empty(1 << 20, dtype=torch.float32, device=torch.device("cuda", local))
t.fill_(float(rank + 1))
hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
print(rank, "ptrs", [hex(p) for p in hdl.buffer_ptrs], "rank", hdl.rank, "world", hdl.world_size, flush=True)
hdl.barrier()
peer = hdl.get_buffer((rank + 1) % world, (1 << 20,), torch.float32)
print(rank, "peer value", float(peer[123].item()), flush=True)
hdl.barrier()
dist.destroy_process_group()
