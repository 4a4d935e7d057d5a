"""The reference's Example/example.py workflow on the drop-in package.

Same sequence of calls and parameters as /Example/example.py (lines 59-147 of the reference): slope, topographic
indexes, downslope, HAND, GFI, ln(hl/H), then HAND is calibrated against the benchmark flood map.  The reference reads
four GeoTIFFs with rasterio (example.py:33-39) and plots with matplotlib; neither is in this image, so the rasters
come from tests/golden/example_inputs.npz (the same bundled 2178 x 1534 example, already nodata-normalised as in
example.py:42-52) and the plots are replaced by a printed summary.  Needs a B200 (no CPU fallback).

    python examples/example.py
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
import descriptools_b200.slope as slope  # noqa: E402
import descriptools_b200.topoindexes as topoindexes  # noqa: E402
from helpers import example_inputs  # noqa: E402


def main():
    ex = example_inputs()
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
    return 0 if same else 1


if __name__ == "__main__":
    sys.exit(main())
