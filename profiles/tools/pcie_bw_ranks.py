"""Copy-only ceiling of the end-to-end leg on N GPUs of one box: every rank streams pinned host memory to its GPU while
copying 0.28x as many bytes back (C2's compression ratio), all ranks at once; aggregate H2D GB/s = what `e2e` could reach
if the kernels were free.  Run under torchrun like bench.py."""
import json, os, time, torch
import torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 2 << 30
m = int(n * 0.28)
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(m, dtype=torch.uint8).pin_memory()
d2 = torch.empty(m, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
flag = torch.zeros(1, device="cuda")
res = {}
for mode in ("h2d", "h2d+d2h"):
    for rep in range(3):
        torch.cuda.synchronize()
        if world > 1:
            dist.all_reduce(flag); torch.cuda.synchronize()
        t = time.perf_counter()
        with torch.cuda.stream(s1):
            for i in range(4): d.copy_(h, non_blocking=True)
        if mode != "h2d":
            with torch.cuda.stream(s2):
                for i in range(4): h2.copy_(d2, non_blocking=True)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t], device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        res[mode] = world * 4 * n / dt.item() / 1e9
if rank == 0:
    print(json.dumps({"n_gpus": world, "aggregate_h2d_GBps_alone": res["h2d"], "aggregate_h2d_GBps_with_0.28x_d2h": res["h2d+d2h"]}))
if world > 1:
    dist.destroy_process_group()
