#!/usr/bin/env python3
"""Concurrent pinned host->device bandwidth of N GPUs of one box (torchrun): what bounds the end-to-end legs when every rank
uploads at once.  Prints one line: per-GPU and aggregate GB/s."""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
a = torch.empty(n, dtype=torch.uint8).pin_memory()
b = torch.empty(n, dtype=torch.uint8, device="cuda")
b.copy_(a, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.time()
for _ in range(4):
    b.copy_(a, non_blocking=True)
torch.cuda.synchronize()
dt = torch.tensor([time.time() - t0], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
if rank == 0:
    per = 4 * n / 1e9 / float(dt.item())
    print(f"concurrent pinned H2D, {world} GPU(s): {per:.1f} GB/s per GPU, {per * world:.1f} GB/s aggregate", flush=True)
if world > 1:
    dist.destroy_process_group()
