"""Times the raster-file paths on a GPU box: host-decoded vs device-decoded read_to_device (per tile size and per
resident-CTA setting of the decode kernel), write_from_device and the file-to-file chain (pipeline.pipeline_files) on
a synthetic DEM.  Prints one JSON line.

    python scripts/time_raster_io.py [size=8192] [out=gpurun_out/raster_io.json] [chain=1]
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
from descriptools_b200 import device, pipeline  # noqa: E402


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
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(REPO, "gpurun_out", "raster_io.json")
    chain = (sys.argv[3] if len(sys.argv) > 3 else "1") == "1"
    dem = device.synth_dem(n, n)  # unconditioned: the files are what is timed here
    dem = (dem * 100).round() / 100  # centimetres, like a DEM product
    raw_mb = dem.numel() * 4 / 1e6
    res = {"rows": n, "cols": n, "raw_MB": raw_mb, "host_threads": os.cpu_count()}
    buf = torch.empty_like(dem)
    with tempfile.TemporaryDirectory() as tmp:
        for tile in (256, 128):
            p = os.path.join(tmp, f"dem{tile}.tif")
            t = best(lambda: rio.write_from_device(p, dem, compress="lzw", tiled=True, blockxsize=tile, blockysize=tile, nodata=-100), 1)
            r = {"file_MB": os.path.getsize(p) / 1e6, "write_from_device_lzw_MBps": raw_mb / t}
            t = best(lambda: rio.read_to_device(p, out=buf, decode="host"), 2)
            assert torch.equal(buf, dem)
            r["read_host_MBps"] = raw_mb / t
            for ctas in (8, 4, 2):
                os.environ["DTB_TIFF_CTAS_PER_SM"] = str(ctas)
                buf.zero_()
                t = best(lambda: rio.read_to_device(p, out=buf, decode="device"), 2)
                assert torch.equal(buf, dem), ctas
                r[f"read_device_ctas{ctas}_MBps"] = raw_mb / t
            os.environ.pop("DTB_TIFF_CTAS_PER_SM")
            res[f"tile{tile}"] = r
        if chain:
            p = os.path.join(tmp, "dem256.tif")
            for mode in ("host", "device"):
                t0 = time.perf_counter()
                pipeline.pipeline_files(p, os.path.join(tmp, "out_" + mode), river_threshold=2000, decode=mode)
                torch.cuda.synchronize()
                res[f"pipeline_files_{mode}_s"] = time.perf_counter() - t0
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as f:
        json.dump(res, f)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
