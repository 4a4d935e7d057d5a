"""Device codec with its tables in global vs shared memory (DTB_TIFF_TABLES): end-to-end rates of read_to_device /
write_from_device and the kernels' own durations (dtb_profile_*), per tile size.  Prints one JSON line.

    python scripts/time_raster_codec_tables.py [size=16384] [out=gpurun_out/raster_codec_tables.json]
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
from descriptools_b200 import _lib, device, pipeline  # noqa: E402


def timed(fn, reps=2):
    """(best wall seconds, {kernel: ms per call}) over `reps` calls"""
    ts = []
    _lib.profile_enable(True)
    _lib.profile_collect()
    for _ in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t)
    prof = {k: round(ms / reps, 3) for k, (ms, n) in _lib.profile_collect().items()}
    _lib.profile_enable(False)
    return min(ts), prof


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(REPO, "gpurun_out", "raster_codec_tables.json")
    dem = device.synth_dem(n, n)
    dem = (dem * 100).round() / 100
    raw_mb = dem.numel() * 4 / 1e6
    res = {"rows": n, "cols": n, "raw_MB": raw_mb}
    buf = torch.empty_like(dem)
    with tempfile.TemporaryDirectory() as tmp:
        for tile in (256, 128):
            kw = dict(compress="lzw", tiled=True, blockxsize=tile, blockysize=tile, nodata=-100)
            for tables in ("global", "shared"):
                os.environ["DTB_TIFF_TABLES"] = tables
                q = os.path.join(tmp, f"dev{tile}{tables}.tif")
                t, prof = timed(lambda: rio.write_from_device(q, dem, encode="device", **kw))
                r = {"write_MBps": raw_mb / t, "write_kernels_ms": prof, "file_MB": os.path.getsize(q) / 1e6}
                buf.zero_()
                t, prof = timed(lambda: rio.read_to_device(q, out=buf, decode="device"))
                assert torch.equal(buf, dem), (tile, tables)
                r.update(read_MBps=raw_mb / t, read_kernels_ms=prof)
                res[f"tile{tile}_{tables}"] = r
        p = os.path.join(tmp, "dev256global.tif")
        for tables in ("global", "shared"):
            os.environ["DTB_TIFF_TABLES"] = tables
            t0 = time.perf_counter()
            pipeline.pipeline_files(p, os.path.join(tmp, "out_" + tables), river_threshold=2000, decode="device", encode="device")
            torch.cuda.synchronize()
            res[f"pipeline_files_{tables}_s"] = time.perf_counter() - t0
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as f:
        json.dump(res, f)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
