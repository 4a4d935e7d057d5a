"""Aggregate host<->device copy bandwidth with every rank copying at once (torchrun): is the e2e figure at N GPUs bound by
the host (shared PCIe uplinks / memory), or by this code?   torchrun --nproc-per-node N scripts/pcie_bw_all.py [GB]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from descriptools_b200 import bands

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
bind = "--no-bind" not in sys.argv
cpus = bands.bind_to_gpu_numa(local) if bind else None
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
gb = float(sys.argv[1]) if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else 4.0
n = int(gb * 1e9 / 4)
host = torch.empty(n, dtype=torch.float32, pin_memory=True)
host.fill_(1.0)
devt = torch.empty(n, dtype=torch.float32, device="cuda")
out = {}
for name, (dst, src) in {"d2h": (host, devt), "h2d": (devt, host)}.items():
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t0) / 3], device="cuda", dtype=torch.float64)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    out[name] = gb / float(dt[0])
if dist.get_rank() == 0:
    w = dist.get_world_size()
    print({"ranks": w, "bound_cores": len(cpus) if cpus else None, "per_gpu_GBps": {k: round(v, 1) for k, v in out.items()},
           "aggregate_GBps": {k: round(v * w, 1) for k, v in out.items()}})
dist.destroy_process_group()
