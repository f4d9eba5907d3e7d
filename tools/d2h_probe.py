#!/usr/bin/env python
"""What bounds the end-to-end arm at N GPUs: every frame ends in a 64 MiB device-to-host copy (get_image_data of a
4096^2 RGBA8 image).  Under torchrun, every rank copies 64 MiB from its GPU into pinned host memory -- first rank 0
alone, then all ranks at once -- and rank 0 prints the per-rank and aggregate GB/s next to the host's NUMA layout.
No canvas code involved: plain cudaMemcpyAsync through torch."""
import json, os, subprocess, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 64 << 20
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
host = torch.empty(n, dtype=torch.uint8, pin_memory=True)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def copies(reps):
    t0 = time.perf_counter()
    for _ in range(reps):
        host.copy_(dev, non_blocking=True)
    torch.cuda.synchronize()
    return n * reps / (time.perf_counter() - t0) / 1e9


copies(5)
barrier()
alone = copies(40) if rank == 0 else 0.0
barrier()
together = copies(40)
barrier()
if world > 1:
    all_rates = [None] * world
    dist.all_gather_object(all_rates, together)
else:
    all_rates = [together]
if rank == 0:
    numa = subprocess.run("lscpu | grep -i -E 'numa|^CPU\\(s\\)|model name'", shell=True, capture_output=True, text=True).stdout
    print(json.dumps({"n_gpus": world, "bytes_per_copy": n, "rank0_alone_gbs": alone, "all_ranks_at_once_gbs": all_rates,
                      "aggregate_gbs": sum(all_rates), "frames_per_s_this_allows": sum(all_rates) * 1e9 / n,
                      "host": numa.strip().splitlines(), "affinity_rank0": sorted(os.sched_getaffinity(0))[:4] + ["...", len(os.sched_getaffinity(0))]}))
if world > 1:
    dist.destroy_process_group()
