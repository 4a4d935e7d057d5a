"""One chain step on a conditioned synthetic DEM -- the command profiled by ncu (see profiles/)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from descriptools_b200 import device, pipeline

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
cols = int(sys.argv[2]) if len(sys.argv) > 2 else rows
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dem = device.conditioned_dem(rows, cols)
torch.cuda.synchronize()
for _ in range(steps):
    res = pipeline.run_device(dem, 12.5, 128000)
torch.cuda.synchronize()
print("ok", rows, cols, int((res["idx"] >= 0).sum()))
