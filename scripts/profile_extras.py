"""The kernels outside the headline chain, once each on a 10k x 10k conditioned DEM -- the command profiled by ncu for
profiles/r2zz_extras_10k_summary.txt (TI + MTI, ln(hl/H), %->radians, calibration counts, downslope)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from descriptools_b200 import device, pipeline, evaluation as ev

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
dem = device.conditioned_dem(n, n)
res = pipeline.run_device(dem, 12.5, 2000)
for _ in range(2):
    rad = device.slope_to_radians(res["slope"])
    ti, mti = device.ti_mti(res["acc"], rad, 12.5, 0.1)
    l = device.ln_hl_H(res["hand"], res["acc"], 0.4, 0.1, 12.5)
    ctr = ev._Counter(res["hand"], False, -100.0, (res["acc"] > 2000).to(torch.int8), "under")
    c = ctr.counts(torch.linspace(0.0, 30.0, 32, dtype=torch.float64).tolist())
    ds = device.downslope(dem, res["d8"], 12.5, 5.0)
torch.cuda.synchronize()
print("ok", n, float(ti[5, 5]), float(l[5, 5]), c[0].tolist(), float(ds[5, 5]))
