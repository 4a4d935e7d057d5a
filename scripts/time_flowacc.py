"""Flow accumulation + fused HAND entry pass alone, per kernel (library profile), on an N x N conditioned DEM.
    python scripts/time_flowacc.py [N=10000] [reps=5]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from descriptools_b200 import _lib, device


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    dem = device.synth_dem(n, n)
    device.fill_depressions(dem)
    _, d8 = device.slope_d8(dem, 12.5)
    for _ in range(2):
        acc = device.flow_accumulation(d8, fuse_hand_threshold=128000)
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    for _ in range(reps):
        acc = device.flow_accumulation(d8, fuse_hand_threshold=128000)
    torch.cuda.synchronize()
    prof = _lib.profile_collect()
    _lib.profile_enable(False)
    print(json.dumps({"n": n, "kernels_ms": {k: round(v[0] / reps, 4) for k, v in prof.items()}, "acc_max": int(acc.max())}))


if __name__ == "__main__":
    main()
