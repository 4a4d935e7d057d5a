"""Times the raster-file paths on a GPU box: host vs device decode (read_to_device) and encode (write_from_device) per
tile size and per resident-CTA setting of the codec kernels, and the file-to-file chain (pipeline.pipeline_files) on a
synthetic DEM.  Prints one JSON line.

    python scripts/time_raster_io.py [size=8192] [out=gpurun_out/raster_io.json] [chain_size=size, 0 = skip]
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


def best(fn, reps=2):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t)
    return min(ts)


def synth(n):
    dem = device.synth_dem(n, n)  # unconditioned: the files are what is timed here
    return (dem * 100).round() / 100  # centimetres, like a DEM product


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(REPO, "gpurun_out", "raster_io.json")
    chain_n = int(sys.argv[3]) if len(sys.argv) > 3 else n
    dem = synth(n)
    raw_mb = dem.numel() * 4 / 1e6
    res = {"rows": n, "cols": n, "raw_MB": raw_mb, "host_threads": os.cpu_count()}
    buf = torch.empty_like(dem)
    with tempfile.TemporaryDirectory() as tmp:
        for tile in (256, 128):
            p, q = os.path.join(tmp, f"dem{tile}.tif"), os.path.join(tmp, f"dev{tile}.tif")
            kw = dict(compress="lzw", tiled=True, blockxsize=tile, blockysize=tile, nodata=-100)
            r = {}
            r["write_host_MBps"] = raw_mb / best(lambda: rio.write_from_device(p, dem, **kw), 1)
            r["write_device_MBps"] = raw_mb / best(lambda: rio.write_from_device(q, dem, encode="device", **kw))
            r["file_MB_host"], r["file_MB_device"] = os.path.getsize(p) / 1e6, os.path.getsize(q) / 1e6
            r["read_host_MBps"] = raw_mb / best(lambda: rio.read_to_device(p, out=buf, decode="host"))
            assert torch.equal(buf, dem)
            for ctas in (8, 4, 2):
                os.environ["DTB_TIFF_CTAS_PER_SM"] = str(ctas)
                buf.zero_()
                t = best(lambda: rio.read_to_device(q, out=buf, decode="device"))
                assert torch.equal(buf, dem), ctas
                r[f"read_device_ctas{ctas}_MBps"] = raw_mb / t
            os.environ.pop("DTB_TIFF_CTAS_PER_SM")
            res[f"tile{tile}"] = r
        if chain_n:
            p = os.path.join(tmp, "chain.tif")
            cdem = dem if chain_n == n else synth(chain_n)
            rio.write_from_device(p, cdem, compress="lzw", tiled=True, blockxsize=256, blockysize=256, nodata=-100)
            res["chain_rows"] = chain_n
            for mode in ("host", "device"):
                t0 = time.perf_counter()
                pipeline.pipeline_files(p, os.path.join(tmp, "out_" + mode), river_threshold=2000, decode=mode, encode=mode)
                torch.cuda.synchronize()
                res[f"pipeline_files_{mode}_s"] = time.perf_counter() - t0
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as f:
        json.dump(res, f)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
