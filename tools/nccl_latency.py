"""All-reduce latency of the gradient-sized buffer (371,907 fp32 = 1.49 MB) over the ranks of one node.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/nccl_latency.py"""
import json
import os
import time

import torch
import torch.distributed as dist

rank, local, world = (int(os.environ.get(k, 0)) for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
out = {}
for n in (4, 371_907, 4_000_000):
    t = torch.ones(n, device="cuda")
    for _ in range(20):
        dist.all_reduce(t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record()
    for _ in range(200):
        dist.all_reduce(t)
    e1.record()
    torch.cuda.synchronize()
    out[f"{n}_floats"] = {"device_us": e0.elapsed_time(e1) * 1000 / 200, "wall_us": (time.perf_counter() - w0) * 1e6 / 200}
if rank == 0:
    print(json.dumps({"n_gpus": world, "allreduce": out}))
dist.destroy_process_group()
