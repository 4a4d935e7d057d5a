"""Deflate / LZW files: host decode (thread team + pinned staging) vs device decode (compressed chunks over PCIe,
one warp per chunk) on a synthetic DEM product, per predictor.  Prints one JSON line.

    python scripts/time_raster_codecs.py [size=8192] [out=gpurun_out/raster_codecs.json]
"""
import json
import os
import sys
import tempfile
import time

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import descriptools_b200.raster as rio  # noqa: E402
from descriptools_b200 import device  # noqa: E402


def best(fn, reps=3):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t)
    return min(ts)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(REPO, "gpurun_out", "raster_codecs.json")
    dem = (device.synth_dem(n, n) * 100).round() / 100  # centimetres, like a DEM product
    raw_mb = dem.numel() * 4 / 1e6
    res = {"rows": n, "cols": n, "raw_MB": raw_mb, "host_threads": os.cpu_count(), "tile": 256}
    buf = torch.empty_like(dem)
    with tempfile.TemporaryDirectory() as tmp:
        for comp in ("lzw", "deflate"):  # (the writer has no PackBits encoder)
            for pred in (1, 3):
                p = os.path.join(tmp, f"{comp}{pred}.tif")
                rio.write_from_device(p, dem, compress=comp, predictor=pred, tiled=True, blockxsize=256, blockysize=256, nodata=-100)
                r = {"file_MB": os.path.getsize(p) / 1e6}
                r["read_host_MBps"] = raw_mb / best(lambda: rio.read_to_device(p, out=buf, decode="host"))
                assert torch.equal(buf, dem)
                buf.zero_()
                r["read_device_MBps"] = raw_mb / best(lambda: rio.read_to_device(p, out=buf, decode="device"))
                assert torch.equal(buf, dem), (comp, pred)
                res[f"{comp}_predictor{pred}"] = r
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as f:
        json.dump(res, f)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
