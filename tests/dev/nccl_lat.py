"""Latency of the tiny exchanges of the sharded protocol (developer tool: torchrun --nproc-per-node N tests/dev/nccl_lat.py)."""
import os
import time

import torch
import torch.distributed as dist

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
world, rank = dist.get_world_size(), dist.get_rank()
for k in (3, 259):
    send = torch.arange(k, dtype=torch.int64, device="cuda")
    recv = torch.empty(k * world, dtype=torch.int64, device="cuda")
    h = torch.empty(k * world, dtype=torch.int64).pin_memory()
    for name, fn in (("all_gather", lambda: dist.all_gather_into_tensor(recv, send)),
                     ("all_reduce", lambda: dist.all_reduce(send))):
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        ts = []
        for _ in range(50):
            t0 = time.perf_counter()
            fn()
            h.copy_(recv, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            ts.append(time.perf_counter() - t0)
        ts.sort()
        if rank == 0:
            print(f"{os.environ.get('TAG', '')} {name} k={k}: median {ts[25] * 1e6:.0f} us  min {ts[0] * 1e6:.0f} us  p90 {ts[45] * 1e6:.0f} us", flush=True)
dist.destroy_process_group()
