#!/usr/bin/env python
"""Aggregate device->host copy bandwidth with all ranks copying at once (run under torchrun):
what the blocking readback of the density can reach on this box, independent of the smoke step."""
import os, time, torch, torch.distributed as dist
r = int(os.environ.get("LOCAL_RANK", 0)); w = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(r)
if w > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", r))
n = 256 * 256 * 256
src = torch.zeros(n, device="cuda"); dst = torch.empty(n, pin_memory=True)
for _ in range(3):
    dst.copy_(src, non_blocking=True)
torch.cuda.synchronize()
if w > 1:
    dist.barrier()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20):
    dst.copy_(src, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
t = torch.tensor([dt], device="cuda")
if w > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if r == 0:
    print(f"D2H {w} ranks x 67 MB x 20: per-rank {n*4*20/t.item()/1e9:.1f} GB/s, aggregate {w*n*4*20/t.item()/1e9:.1f} GB/s")
if w > 1:
    dist.destroy_process_group()
