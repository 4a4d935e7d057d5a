"""Per-kernel CUDA-event timings of the chain (library profiling hooks) at a given size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from descriptools_b200 import device, _lib, pipeline

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
cols = int(sys.argv[2]) if len(sys.argv) > 2 else rows
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
thr = int(sys.argv[4]) if len(sys.argv) > 4 else 128000
dem = device.conditioned_dem(rows, cols)
for _ in range(2):
    pipeline.run_device(dem, 12.5, thr)
torch.cuda.synchronize()
_lib.profile_enable(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    res = pipeline.run_device(dem, 12.5, thr)
e1.record()
torch.cuda.synchronize()
tot = e0.elapsed_time(e1) / reps
k = _lib.profile_collect()
for name, (ms, n) in k.items():
    print(f"{ms/reps:9.3f} ms  x{n//reps:<3d} {name}")
print(f"{sum(v[0] for v in k.values())/reps:9.3f} ms  kernels; {tot:.3f} ms per step; {rows*cols/tot/1e3:.1f} Mcells/s")
