"""Host <-> device copy bandwidth with all ranks of a node copying at once (pinned memory): H2D
alone, D2H alone, and both concurrently on two streams - the ceiling of the end-to-end path when N
processes share the host.  Launch with torchrun like bench.py; prints one line per rank and mode."""
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 29
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.zeros(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for mode in ("h2d", "d2h", "both"):
    for rep in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if rep == 2:
            print(f"rank {rank}/{world} {mode}: {n / dt * 1e-9:.1f} GB/s per direction (512 MiB, {dt * 1e3:.1f} ms)", flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
