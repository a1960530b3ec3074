"""2-GPU probe: does torch.distributed._symmetric_memory work on this box (peer pointers, device barrier)?
torchrun --nproc-per-node 2 tools/micro/symm_probe.py"""
import os
import time
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
t = symm_mem.empty(1024 * 1024, dtype=torch.float32, device=dev)
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, "ptrs", [hex(p) for p in hdl.buffer_ptrs], "multicast_ptr", hex(hdl.multicast_ptr) if hdl.multicast_ptr else None, flush=True)
t.fill_(float(rank + 1))
hdl.barrier()
peer = hdl.get_buffer((rank + 1) % world, (1024 * 1024,), torch.float32)
print(rank, "peer value", float(peer[0]), flush=True)
# write into the peer's buffer, barrier, read own
peer[:16] = 100.0 + rank
hdl.barrier()
torch.cuda.synchronize()
print(rank, "own after peer write", t[:2].tolist(), flush=True)
# barrier latency
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(100):
    hdl.barrier()
b.record()
torch.cuda.synchronize()
print(rank, "barrier us", a.elapsed_time(b) * 10, flush=True)
dist.destroy_process_group()
