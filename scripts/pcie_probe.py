#!/usr/bin/env python
"""Plain pinned-memory copy probe, one process per GPU (torchrun): what the host gives N GPUs at
once, with nothing of this repo in the way.  Every rank copies a 1 GiB device buffer to pinned
host memory (D2H) and back (H2D) ITERS times with cudaMemcpyAsync on its own stream; ranks start
together; per-rank and aggregate GB/s are printed by rank 0.
   python -m torch.distributed.run --nproc-per-node N scripts/pcie_probe.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wssdl_bus_b200.pipeline import bind_to_gpu_numa_node  # noqa: E402

ITERS, NBYTES = 8, 1 << 30
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
node = bind_to_gpu_numa_node(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.empty(NBYTES, dtype=torch.uint8, device="cuda")
host = torch.empty(NBYTES, dtype=torch.uint8).pin_memory()
res = {}
s2 = torch.cuda.Stream()
host2 = torch.empty(NBYTES, dtype=torch.uint8).pin_memory()
dev2 = torch.empty(NBYTES, dtype=torch.uint8, device="cuda")


def one(name):
    if name == "d2h":
        host.copy_(dev, non_blocking=True)
    elif name == "h2d":
        dev.copy_(host, non_blocking=True)
    else:                                       # full duplex: D2H here, H2D on a second stream
        host.copy_(dev, non_blocking=True)
        with torch.cuda.stream(s2):
            dev2.copy_(host2, non_blocking=True)


for name in ("d2h", "h2d", "both"):
    for it in range(2):
        one(name)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for it in range(ITERS):
        one(name)
    torch.cuda.current_stream().wait_stream(s2)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    res[name] = ITERS * NBYTES * (2 if name == "both" else 1) / ms / 1e6
t = torch.tensor([res["d2h"], res["h2d"], res["both"]], device="cuda", dtype=torch.float64)
if world > 1:
    allr = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allr, t)
else:
    allr = [t]
if rank == 0:
    rows = [[float(v) for v in r.tolist()] for r in allr]
    print(json.dumps({"ranks": world, "numa_node_rank0": node, "host_cpus": os.cpu_count(),
                      "per_rank_gbs_d2h": [round(r[0], 1) for r in rows],
                      "per_rank_gbs_h2d": [round(r[1], 1) for r in rows],
                      "aggregate_gbs_d2h": round(sum(r[0] for r in rows), 1),
                      "aggregate_gbs_h2d": round(sum(r[1] for r in rows), 1),
                      "aggregate_gbs_duplex": round(sum(r[2] for r in rows), 1)}))
if world > 1:
    dist.destroy_process_group()
