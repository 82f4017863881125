#!/usr/bin/env python3
"""Host-side ceiling of the end-to-end leg at N GPUs: every rank copies pinned host memory to its GPU and back at the same
time (the byte counts of bench.py's e2e step: 32 4K frames up + their payloads, and the mirror image down), all ranks
started together.  Aggregate GB/s over all ranks = total bytes / slowest rank's time.  Run under torchrun; one JSON line.
   torchrun --nproc-per-node N profiles/pcie_ceiling.py >> profiles/r02_pcie_ceiling.jsonl"""
import json, os, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
UP = 32 * 3840 * 2160 * 3 // 2 + 32 * 2813348   # frames + payloads (what compress_batch_host + decompress_batch_host upload per step)
DOWN = UP
d1 = torch.empty(UP, dtype=torch.uint8, device=dev); d2 = torch.empty(DOWN, dtype=torch.uint8, device=dev)
h1 = torch.empty(UP, dtype=torch.uint8).pin_memory(); h2 = torch.empty(DOWN, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def step():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
def run(reps):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): step()
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
run(3)
reps = 12
t = run(reps)
if rank == 0:
    total = world * reps * (UP + DOWN)
    print(json.dumps({"n_gpus": world, "aggregate_two_way_GBps": round(total / t / 1e9, 1), "per_gpu_two_way_GBps": round(total / t / 1e9 / world, 1),
                      "ms_per_step": round(1e3 * t / reps, 3), "bytes_per_step_per_gpu": UP + DOWN,
                      "e2e_ceiling_Mpixel_s": round(world * 32 * 3840 * 2160 / (t / reps) / 1e6, 1), "host_cpus": os.cpu_count()}), flush=True)
if world > 1: dist.destroy_process_group()
