"""Pinned host<->device copy bandwidth of this box (context for the e2e number).

    python tools/h2d_bw.py                                   one GPU
    python -m torch.distributed.run --nproc-per-node 8 ... tools/h2d_bw.py     all GPUs AT ONCE

Under torchrun every rank copies to its own GPU at the same time (barrier first), which is what
the e2e arm of `bench.py --gpus N` does: the per-rank rate then shows what the host side of the
box (memory bandwidth, PCIe switch uplinks shared by GPU pairs, IOMMU) can feed concurrently."""
import json
import os
import time

import torch

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 168 * 1024 * 1024
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
out = {"rank": rank, "world": world, "affinity": len(os.sched_getaffinity(0))}
for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    out[name + "_gbs"] = n / dt / 1e9
print(json.dumps(out), flush=True)
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
