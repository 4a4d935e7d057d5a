"""ACTIVE entry nodes left after each hand_node_jump launch (the counters at the head of the HAND workspace).
    python scripts/node_jump_active.py [N=20000]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from descriptools_b200 import device, pipeline

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
dem = device.conditioned_dem(n, n)
res = pipeline.run_device(dem, 12.5, 128000)
torch.cuda.synchronize()
for key, buf in device.workspace.buf.items():
    if key[1] == "hand":
        head = buf[:256].view(torch.int32).cpu().tolist()
        print("nodes (slots):", (n + 63) // 64 * ((n + 63) // 64) * 256, "active after launch r:", head[:12])
