"""Print the cells where the device slope / D8 differ from the oracle on the near-tie test raster (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from descriptools_b200 import device

def case(px):
    rng = np.random.default_rng(3)
    rows, cols = 96, 128
    dem = np.full((rows, cols), 500.0, np.float32)
    s2 = np.sqrt(2.0)
    for r in range(1, rows - 1, 3):
        for c in range(1, cols - 1, 3):
            a = np.float32(rng.uniform(0.5, 30))
            eps = np.float32(rng.choice([0, 1, -1, 2, -2, 50, -50])) * np.spacing(a)
            dem[r, c + 1] = np.float32(500.0) - a
            dem[r + 1, c + 1] = np.float32(500.0) - np.float32(np.float32(a * s2) + eps)
    return dem

for px in [12.5, 30.0, 1.0, 0.3, 7.77]:
    dem = case(px)
    s_ref, d_ref = oracle.slope_d8(dem, px)
    s, d = device.slope_d8(torch.from_numpy(dem).cuda(), px)
    s, d = s.cpu().numpy(), d.cpu().numpy()
    bad = np.argwhere((s != s_ref) | (d != d_ref))
    print("px", px, "mismatches", len(bad))
    for r, c in bad[:6]:
        print("  cell", r, c, "got", s[r, c], d[r, c], "ref", s_ref[r, c], d_ref[r, c])
        print(np.array2string(dem[max(r-1,0):r+2, max(c-1,0):c+2], precision=9, floatmode="unique"))
