"""Per-stage CUDA-event timings of the chain at a given size (development aid, not the bench)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from descriptools_b200 import device, _lib

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
cols = int(sys.argv[2]) if len(sys.argv) > 2 else rows
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
thr = int(sys.argv[4]) if len(sys.argv) > 4 else 128000
dem = device.conditioned_dem(rows, cols)
torch.cuda.synchronize()
it = torch.int32 if rows * cols < 2**31 else torch.int64
ev = lambda: torch.cuda.Event(enable_timing=True)
for rep in range(reps):
    e = [ev() for _ in range(4)]
    e[0].record()
    slope, d8 = device.slope_d8(dem, 12.5)
    e[1].record()
    acc = device.flow_accumulation(d8, dtype=it)
    e[2].record()
    out = device.hand(d8, dem, 12.5, acc=acc, river_threshold=thr, gfi_params=(0.4, 0.1, 12.5), idx_dtype=it)
    e[3].record()
    torch.cuda.synchronize()
    t = [e[i].elapsed_time(e[i + 1]) for i in range(3)]
    print(f"rep {rep}: slope_d8 {t[0]:.3f} ms  flowacc {t[1]:.3f} ms  hand_gfi {t[2]:.3f} ms  total {sum(t):.3f} ms  "
          f"{rows*cols/sum(t)/1e3:.1f} Mcells/s", flush=True)
    del slope, acc, out
print("max acc", "n/a", "launches", _lib.launch_count())
