import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from descriptools_b200 import device
rows = int(sys.argv[1]); cols = int(sys.argv[2])
dem = device.conditioned_dem(rows, cols)
slope, d8 = device.slope_d8(dem, 12.5)
for _ in range(2):
    acc = device.flow_accumulation(d8)
torch.cuda.synchronize()
ws = device.workspace.buf[0]
print("counters", ws[:32].view(torch.int64).tolist(), "acc max", int(acc.max()))
