"""dtb_eval_counts: kernel time (library profile) vs time of the whole call, 10k x 10k f32 descriptor, 32 thresholds."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from descriptools_b200 import _lib, device, pipeline, evaluation as ev

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
dem = device.conditioned_dem(n, n)
res = pipeline.run_device(dem, 12.5, 2000)
desc = res["hand"]
bench_map = (res["acc"] > 2000).to(torch.int8)
ths = torch.linspace(0.0, 30.0, 32, dtype=torch.float64).tolist()
out = {}
for under in ("under", "over"):
    ctr = ev._Counter(desc, False, -100.0, bench_map, under)
    for _ in range(2):
        ctr.counts(ths)
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        c = ctr.counts(ths)
    e1.record()
    torch.cuda.synchronize()
    prof = _lib.profile_collect()
    _lib.profile_enable(False)
    out[under] = {"call_ms": e0.elapsed_time(e1) / 5, "kernels_ms": {k: v[0] / v[1] for k, v in prof.items()}, "counts_row0": c[0].tolist()}
print(json.dumps(out))
