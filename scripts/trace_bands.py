"""Where a band step spends its time (rank 0): CUDA events around summary / gather / solve / scatter / finish."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from descriptools_b200 import bands, device

local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
rows = cols = 40000
runner = bands.BandRunner(rows, cols, 12.5, 128000, 0.4, 0.1)
band = runner.bands[0]
band.dem.copy_(device.synth_dem(band.rows, cols, band.r0))   # unfilled DEM is fine for timing the plumbing
torch.cuda.synchronize()
x = runner.x
ev = lambda: torch.cuda.Event(enable_timing=True)
def traced_solve(per_band, solver, rounds=bands.ROUNDS):
    (mine,) = per_band
    e = [ev() for _ in range(5)]
    e[0].record()
    mine = mine.contiguous()
    gathered = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, gathered, dst=0)
    e[1].record()
    out = torch.empty(x._out_shape(mine, solver), dtype=mine.dtype, device=mine.device)
    parts = None
    if rank == 0:
        res, flag = x.graphed(solver, torch.stack(gathered, 0), rounds)
        x.flags.append(flag.clone())
        parts = [res[i].contiguous() for i in range(world)]
    e[2].record()
    dist.scatter(out, parts, src=0)
    e[3].record()
    trace.append((solver.__name__, e))
    return [out]
x.solve = traced_solve
for it in range(6):
    trace = []
    evs = [ev() for _ in range(4)]
    t0 = time.perf_counter()
    runner.step(evs)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    if it >= 3:
        msg = f"rank {rank} it {it}: wall {wall:.2f} stages " + " ".join(f"{evs[i].elapsed_time(evs[i+1]):.2f}" for i in range(3))
        for name, e in trace:
            msg += f" | {name[6:13]}: gather {e[0].elapsed_time(e[1]):.2f} solve {e[1].elapsed_time(e[2]):.2f} scatter {e[2].elapsed_time(e[3]):.2f}"
        print(msg, flush=True)
dist.destroy_process_group()
