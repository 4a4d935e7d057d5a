"""Fused slope + D8 stencil alone (BASELINE configs[1]): CUDA-event time per launch on three 10k x 10k DEMs --
conditioned (depression-filled), unconditioned (pits), and conditioned with nodata holes -- plus a bit-exact
check of a 2048 x 2048 crop of each against the oracle.

    python scripts/time_stencil.py [N=10000] [reps=20]      (env DTB_STENCIL_* select A/B variants)
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import oracle
from descriptools_b200 import device

PX = 12.5


def holes(dem):
    """nodata blobs (~8 % of the raster): discs of radius 40..400 cells on a fixed grid of centres"""
    n, m = dem.shape
    out = dem.clone()
    g = torch.Generator(device="cpu").manual_seed(7)
    k = max(4, n // 1000)
    cy = (torch.rand(k * k, generator=g) * n).long()
    cx = (torch.rand(k * k, generator=g) * m).long()
    rr = (40 + torch.rand(k * k, generator=g) * 360).long()
    for y, x, r in zip(cy.tolist(), cx.tolist(), rr.tolist()):
        y0, y1, x0, x1 = max(0, y - r), min(n, y + r + 1), max(0, x - r), min(m, x + r + 1)
        yy = torch.arange(y0, y1, device=dem.device).view(-1, 1) - y
        xx = torch.arange(x0, x1, device=dem.device).view(1, -1) - x
        out[y0:y1, x0:x1][(yy * yy + xx * xx) <= r * r] = -100.0
    return out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    raw = device.synth_dem(n, n)
    cond = raw.clone()
    device.fill_depressions(cond)
    cases = {"conditioned": cond, "unconditioned": raw, "nodata_holes": holes(cond)}
    slope = torch.empty((n, n), dtype=torch.float32, device="cuda")
    d8 = torch.empty((n, n), dtype=torch.uint8, device="cuda")
    res = {"n": n, "variant": {k: v for k, v in os.environ.items() if k.startswith("DTB_STENCIL")}}
    for name, dem in cases.items():
        def run():
            device.check(device.lib.dtb_slope_d8(dem.data_ptr(), 0, n, n, 0, n, PX, slope.data_ptr(), d8.data_ptr(),
                                                 torch.cuda.current_stream().cuda_stream), "dtb_slope_d8")
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        # parity on a crop (own halo: the crop is a raster of its own)
        c = min(n, 2048)
        r0 = (n - c) // 2
        crop = dem[r0:r0 + c, r0:r0 + c].contiguous()
        s_c, d_c = device.slope_d8(crop, PX)
        s_ref, d_ref = oracle.slope_d8(crop.cpu().numpy(), PX)
        ok_s = bool(np.array_equal(s_c.cpu().numpy(), s_ref))
        ok_d = bool(np.array_equal(d_c.cpu().numpy(), d_ref))
        pits = float(((d8 == 0) & (dem > -100)).float().mean())
        res[name] = {"ms": round(ms, 4), "gbs": round(9 * n * n / ms / 1e6, 1), "frac_6560": round(9 * n * n / ms / 1e6 / 6560, 3),
                     "slope_exact": ok_s, "d8_exact": ok_d, "nodata_frac": round(float((dem <= -100).float().mean()), 4),
                     "code0_valid_frac": round(pits, 5)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
