"""Where a band step spends its time (rank 0): CUDA events around the halo exchange, the per-band calls and the two
all-gather + boundary-solve rounds.   torchrun --nproc-per-node N scripts/trace_bands.py [ROWS] [COLS]
(ROWS = N * 5000 gives every rank the band it has in the 8-GPU run of the 40 000 x 40 000 workload)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from descriptools_b200 import bands, device

local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 40000
runner = bands.BandRunner(rows, cols, 12.5, 128000, 0.4, 0.1)
band = runner.bands[0]
band.dem.copy_(device.synth_dem(band.rows, cols, band.r0))   # unfilled DEM is fine for timing the plumbing
torch.cuda.synchronize()
x = runner.x
ev = lambda: torch.cuda.Event(enable_timing=True)
marks = []

def mark(name):
    e = ev(); e.record(); marks.append((name, e))

orig_halo, orig_solve = x.halo, x.solve
def halo(items):
    mark("halo<"); orig_halo(items); mark("halo>")
def solve(per_band, solver, rounds=bands.ROUNDS):
    mark(solver.__name__[6:13] + "<")
    (mine,) = per_band
    mine = mine.contiguous()
    everyone = torch.empty((world,) + tuple(mine.shape), dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(everyone.view(-1), mine.view(-1))
    mark("gathered")
    res, flag = x.graphed(solver, everyone, rounds)
    x.flags.append(flag.clone())
    out = [res[rank].clone()]
    mark("solved")
    return out
x.halo, x.solve = halo, solve
for name in ("slope_d8", "flowacc_summary", "flowacc_finish", "hand_summary", "hand_finish"):
    f = getattr(band, name)
    def wrap(f=f, name=name):
        def g(*a, **k):
            mark(name + "<"); r = f(*a, **k); mark(name + ">"); return r
        return g
    setattr(band, name, wrap())
for it in range(6):
    marks.clear()
    t0 = time.perf_counter()
    mark("start")
    runner.step(check=False)
    mark("end")
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    if it >= 4 and rank == 0:
        t = [(n, marks[0][1].elapsed_time(e)) for n, e in marks]
        print(f"it {it}: wall {wall:.2f} ms; " + " ".join(f"{n}@{v:.3f}" for n, v in t), flush=True)
        print("   deltas: " + " ".join(f"{t[i+1][0]}:{t[i+1][1]-t[i][1]:.3f}" for i in range(len(t) - 1)), flush=True)
runner.check_deferred()
dist.destroy_process_group()
