"""How long does the device depression fill take at a given size?  python scripts/time_fill.py ROWS [COLS]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from descriptools_b200 import device

rows = int(sys.argv[1]); cols = int(sys.argv[2]) if len(sys.argv) > 2 else rows
t0 = time.perf_counter()
dem = device.synth_dem(rows, cols)
torch.cuda.synchronize()
t1 = time.perf_counter()
it = device.fill_depressions(dem)
torch.cuda.synchronize()
t2 = time.perf_counter()
print({"rows": rows, "cols": cols, "synth_s": round(t1 - t0, 2), "fill_s": round(t2 - t1, 2), "passes": it,
       "peak_gb": round(torch.cuda.max_memory_allocated() / 1e9, 1)})
