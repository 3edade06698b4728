"""Concurrent host -> device bandwidth of the box: every rank copies the same number of bytes from its own pinned buffer
to its own GPU at the same time (what the e2e step of bench.py does with psi).  torchrun or plain python (1 rank).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/h2d_concurrent.py [MiB]"""
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
host = torch.empty(mib << 20, dtype=torch.uint8).pin_memory()
dev = torch.empty_like(host, device="cuda")
for active in ([world] if world == 1 else [1, 2, 4, world]):
    for _ in range(2):
        dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    reps = 8
    if rank < active:
        for _ in range(reps):
            dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = torch.tensor([reps * (mib << 20) / dt / 1e9 if rank < active else 0.0], device="cuda", dtype=torch.float64)
    if world > 1:
        allv = [torch.zeros_like(gbs) for _ in range(world)]
        dist.all_gather(allv, gbs)
        dist.barrier()
    else:
        allv = [gbs]
    if rank == 0:
        vals = [float(x[0]) for x in allv][:active]
        print(f"{active} rank(s) copying {mib} MiB each at the same time: per rank {', '.join('%.1f' % v for v in vals)} GB/s; "
              f"aggregate {sum(vals):.1f} GB/s", flush=True)
if world > 1:
    dist.destroy_process_group()
