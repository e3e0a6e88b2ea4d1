"""all_to_all_single bandwidth probe (development aid)."""
import os, sys, time, torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w = dist.get_world_size(); r = dist.get_rank()
for mb in (4, 64, 400):
    n = mb * (1 << 20) // 4
    send = torch.ones(n * w, dtype=torch.int32, device="cuda"); recv = torch.empty_like(send)
    for _ in range(3): dist.all_to_all_single(recv, send)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): dist.all_to_all_single(recv, send)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    if r == 0: print("a2a %d MB per peer: %.3f ms -> %.1f GB/s out per rank (excluding self)" % (mb, dt * 1e3, mb * (w - 1) / 1024 / dt), flush=True)
# symmetric memory availability
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(1 << 20, dtype=torch.int32, device="cuda")
    h = symm.rendezvous(t, dist.group.WORLD)
    if r == 0: print("symm ok", type(h), len(h.buffer_ptrs), [hex(p) for p in h.buffer_ptrs][:2], flush=True)
    h.barrier()
except Exception as e:
    if r == 0: print("symm FAILED", repr(e)[:300], flush=True)
dist.destroy_process_group()
