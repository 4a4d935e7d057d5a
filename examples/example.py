"""The reference's Example/example.py workflow on the drop-in package.

Same sequence of calls and parameters as /Example/example.py (lines 59-147 of the reference): slope, topographic
indexes, downslope, HAND, GFI, ln(hl/H), then HAND is calibrated against the benchmark flood map and the class map is
exported as a GeoTIFF with the DEM's georeferencing (example.py:201-217).  The reference reads and writes GeoTIFFs with
rasterio (example.py:33-39, :106, :216) and plots with matplotlib; neither is in this image: `descriptools_b200.raster`
stands in for rasterio (same calls), the plots are replaced by a printed summary.  Needs a B200 (no CPU fallback).

    python examples/example.py <reference>/Example       # the four input GeoTIFFs, as example.py:33-52 reads them
    python examples/example.py                            # the same rasters from tests/golden/example_inputs.npz
"""
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

import descriptools_b200.downslope as downslope  # noqa: E402
import descriptools_b200.evaluation as evaluation  # noqa: E402
import descriptools_b200.flowhand as flowhand  # noqa: E402
import descriptools_b200.gfi as gfi  # noqa: E402
import descriptools_b200.raster as rio  # noqa: E402
import descriptools_b200.slope as slope  # noqa: E402
import descriptools_b200.topoindexes as topoindexes  # noqa: E402
from helpers import example_inputs  # noqa: E402


def read_inputs(folder):
    """example.py:33-52 and :106, verbatim apart from the module behind `rio`"""
    with np.errstate(invalid="ignore"):
        dem = rio.open(os.path.join(folder, "input/12_dem.tif")).read(1).astype("int16")
        fdr = rio.open(os.path.join(folder, "input/12_fdr.tif")).read(1)
        fac = rio.open(os.path.join(folder, "input/12_fac.tif")).read(1).astype("int")
    dem = np.where(dem == dem[0, 0], -100, dem)
    fac = np.where(fac == fac[0, 0], -100, fac)
    river = np.where(fac > 128000, 1, 0).astype("int8")
    flood = rio.open(os.path.join(folder, "input/WB_12_100y.tif")).read(1).astype("int8")
    return dem, fdr, fac, river, flood


def main(folder=None, out_file="hand_class.tif"):
    ex = example_inputs()
    if folder:
        dem, fdr, fac, river, flood = read_inputs(folder)
    else:
        dem, fdr, fac, river, flood = ex["dem"], ex["fdr"], ex["fac"], ex["river"], ex["flood"].copy()
    px = 12.5
    t0 = time.perf_counter()
    sl = slope.sloper(dem, px)                                                    # example.py:59
    sl_rad = np.where(sl == -100, -100, np.arctan(sl / 100)).astype("float32")    # example.py:63-64
    ti, mti = topoindexes.topographic_index(fac, sl_rad, px, 0.1)                 # example.py:69
    down = downslope.downsloper(dem, fdr, px, 5)                                  # example.py:74
    fdist, indices, hand = flowhand.flow_hand_index(dem, fdr, river, px)          # example.py:82
    gfi_map = gfi.gfi_calculator(hand, fac, indices, 0.4, 0.1, px)                # example.py:87
    lnhlh = gfi.ln_hl_H_calculator(hand, fac, 0.4, 0.1, px)                       # example.py:91
    elements = np.unique(hand)                                                    # example.py:113-115
    mx, mn = elements[-1], elements[1]
    desc = evaluation.minMaxScale(hand, mn, mx, -100)                             # example.py:121
    th = evaluation.calibration(desc, flood, "under")                             # example.py:136
    binary = evaluation.binary_map(desc, th, "under")                             # example.py:139
    c, f, class_map = evaluation.avaliacao(binary, flood)                         # example.py:147
    dt = time.perf_counter() - t0
    valid = dem != -100
    print(f"rasters {dem.shape[0]} x {dem.shape[1]}, {int(valid.sum())} valid cells, {int(river.sum())} river cells, {dt:.2f} s")
    for name, a in (("slope %", sl), ("TI", ti), ("MTI", mti), ("downslope", down), ("flow distance", fdist), ("HAND", hand),
                    ("GFI", gfi_map), ("ln(hl/H)", lnhlh)):
        v = a[valid & (a != -100)]
        print(f"  {name:14s} min {v.min():10.4f}  mean {v.mean():10.4f}  max {v.max():10.4f}")
    print(f"HAND calibration: threshold {th}, correctness {c:.6f}, fit {f:.6f}")
    counts = [int((class_map == k).sum()) for k in range(4)]
    print("class map counts (tn, fp, fn, tp):", counts)
    same = np.array_equal(class_map.astype(np.uint8), ex["hand_class"])
    print("matches the reference's Example/output/hand_class.tif:", same)
    if folder:
        meta = rio.open(os.path.join(folder, "input/12_dem.tif")).meta                # example.py:202
        meta.update(dtype=rio.uint8)                                                  # example.py:205
        meta.update(nodata=0)                                                         # example.py:206
        class_map = class_map.astype("uint8")
        class_map = class_map.reshape(1, len(class_map), len(class_map[0]))           # example.py:211-212
        with rio.open(out_file, "w", **meta) as dist:                                 # example.py:216-217
            dist.write(class_map.astype(rio.uint8))
        print("wrote", out_file, os.path.getsize(out_file), "bytes")
    return 0 if same else 1


if __name__ == "__main__":
    sys.exit(main(*sys.argv[1:3]))
