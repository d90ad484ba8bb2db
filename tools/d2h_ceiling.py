"""The box's device-to-host ceiling: every rank copies a 1 GiB device buffer into its own pinned host buffer, all ranks at
once, with plain asynchronous copies (torch -> cudaMemcpyAsync) -- nothing of this repository on the path. Run with
N = 1 and N = 8 to see what aggregate D2H rate the host side sustains; bench.py's complete copy-out (e2e.full) cannot
exceed it.

  python tools/d2h_ceiling.py
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/d2h_ceiling.py
"""
import json
import os
import time

import torch

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
try:
    import pynvml
    pynvml.nvmlInit()
    pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
except Exception:
    pass
n = 1 << 30
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
for _ in range(2):
    host.copy_(dev, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
reps = 10
t0 = time.perf_counter()
for _ in range(reps):
    host.copy_(dev, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
sec = time.perf_counter() - t0
mine = torch.tensor([reps * n / sec / 1e9], dtype=torch.float64, device="cuda")
if world > 1:
    allr = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allr, mine)
    rates = [float(x.item()) for x in allr]
else:
    rates = [float(mine.item())]
if rank == 0:
    print(json.dumps({"gpus": world, "per_gpu_GBps": [round(r, 2) for r in rates], "aggregate_GBps": round(sum(rates), 2),
                      "bytes_per_copy": n, "copies": reps, "cpus_bound": len(os.sched_getaffinity(0))}))
if world > 1:
    dist.destroy_process_group()
