import os, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import torch.distributed._symmetric_memory as symm_mem
t = symm_mem.empty(1 << 16, dtype=torch.float64, device="cuda")
h = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
print(rank, "buffer_ptrs", [hex(p) for p in h.buffer_ptrs], "signal_pad_ptrs", [hex(p) for p in h.signal_pad_ptrs], "size", h.buffer_size, flush=True)
t.fill_(rank + 1)
h.barrier()
peer = h.get_buffer((rank + 1) % world, (8,), torch.float64)
print(rank, "peer view", peer.tolist()[:2], flush=True)
# timing of NCCL small all-reduce for reference
x = torch.ones(4000, dtype=torch.float64, device="cuda")
for _ in range(20): dist.all_reduce(x)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(200): dist.all_reduce(x)
ev[1].record(); torch.cuda.synchronize()
print(rank, "nccl all_reduce 32KB f64: %.1f us" % (ev[0].elapsed_time(ev[1]) * 1e3 / 200), flush=True)
dist.destroy_process_group()
