#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the REFERENCE itself.

Run in the build container only (needs /root/reference, numba, PIL):

    python tests/golden/make_golden.py            # everything
    python tests/golden/make_golden.py --only cudasim

Outputs (committed; the GPU box has no /root/reference):
  example_inputs.npz     the bundled Example/input rasters after example.py:33-52 glue
                         (int16 DEM, u8 FDR, int32 FAC with -100 nodata, packed flood map)
                         and the reference's only golden output, hand_class.tif (KAT-1).
  example_cpujit.npz     reference compiled CPU-jit twins on the full example:
                         sha256 of every output + a strided sample + summary stats.
  cudasim_*.npz          the UNMODIFIED @cuda.jit kernels, driven through the reference's
                         public entry points under NUMBA_ENABLE_CUDASIM=1, on small rasters.

CUDASIM typing note: the simulator runs kernel bodies on NumPy
scalars, so under NEP 50 `f32 (op) python-float` stays f32 whereas compiled Numba
promotes to f64.  To make the simulator reproduce COMPILED semantics exactly we pass
scalars as np.float64 and pass f32-valued `hand` / `slope` inputs of the pointwise
kernels as f64 arrays (so `x + 0.01` is evaluated in f64, as compiled code does).
"""
import argparse
import hashlib
import os
import sys
import warnings

os.environ["NUMBA_ENABLE_CUDASIM"] = "1"  # must precede the numba import
warnings.simplefilter("ignore")

import numpy as np  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(0, REPO)

import oracle  # noqa: E402  (only used to synthesise D8/acc inputs the reference cannot make)

PX = np.float64(12.5)


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_example():
    from PIL import Image

    Image.MAX_IMAGE_PIXELS = None
    d = os.path.join(REF, "Example", "input")
    rd = lambda p: np.array(Image.open(p))
    dem = rd(os.path.join(d, "12_dem.tif")).astype("int16")  # example.py:33
    fdr = rd(os.path.join(d, "12_fdr.tif"))  # example.py:36
    fac = rd(os.path.join(d, "12_fac.tif")).astype("int")  # example.py:39
    dem = np.where(dem == dem[0, 0], -100, dem)  # example.py:42
    fac = np.where(fac == fac[0, 0], -100, fac)  # example.py:43
    flood = rd(os.path.join(d, "WB_12_100y.tif")).astype("int8")  # example.py:106
    hand_class = rd(os.path.join(REF, "Example", "output", "hand_class.tif"))
    return dem.astype(np.int16), fdr.astype(np.uint8), fac.astype(np.int64), flood, hand_class


def make_example_inputs():
    dem, fdr, fac, flood, hand_class = load_example()
    assert fac.max() < 2**31
    np.savez_compressed(
        os.path.join(HERE, "example_inputs.npz"),
        dem=dem, fdr=fdr, fac=fac.astype(np.int32),
        flood_bits=np.packbits(flood.astype(np.uint8)), hand_class=hand_class.astype(np.uint8),
        shape=np.array(dem.shape),
    )
    print("example_inputs.npz", dem.shape)


def make_example_cpujit():
    import descriptools.slope as S
    import descriptools.flowhand as F
    import descriptools.downslope as D
    import descriptools.gfi as G

    dem, fdr, fac, _, _ = load_example()
    river = np.where(fac > 128000, 1, 0).astype("int8")  # example.py:52
    out = {}
    slope = S.slope_sequential_jit(dem, 12.5).astype("float32")
    fdist, idx = F.fdist_indexes_sequential_jit(fdr, river, 12.5)
    idx = np.ascontiguousarray(idx.astype("int64"))
    hand = np.ascontiguousarray(F.hand_calculator(dem, idx))
    down = D.downslope_sequential_jit(dem, fdr, 12.5, 5)
    gfi = G.geomorphic_flood_index_sequential_jit(hand, np.ascontiguousarray(fac), idx, 0.4, 0.1, 12.5)
    lnh = G.ln_hl_H_sequential_jit(hand, np.ascontiguousarray(fac), 0.4, 0.1, 12.5)
    for name, a in dict(slope=slope, fdist=fdist, idx=idx, hand=hand, downslope=down, gfi=gfi, lnhlh=lnh).items():
        a = np.ascontiguousarray(a)
        out[name + "_sha256"] = np.array(sha(a))
        out[name + "_dtype"] = np.array(str(a.dtype))
        out[name + "_sample"] = a.reshape(-1)[::37].copy()
        print(name, a.dtype, sha(a)[:16])
    np.savez_compressed(os.path.join(HERE, "example_cpujit.npz"), **out)


def run_reference_gpu_entry_points(dem, fdr, fac, river, n_top, delta, n_gfi, b):
    """example.py:59-91 through the reference's public (CUDASIM) entry points."""
    import descriptools.slope as S
    import descriptools.flowhand as F
    import descriptools.downslope as D
    import descriptools.gfi as G
    import descriptools.topoindexes as T

    r = {}
    sl = S.sloper(dem, PX).astype("float32")  # example.py:59
    r["slope"] = sl
    sl_rad = np.arctan(sl / 100).astype("float32")  # example.py:63
    sl_rad = np.where(dem == -100, -100, sl_rad).astype("float32")  # example.py:64
    r["slope_rad"] = sl_rad
    ti, mti = T.topographic_index(fac, sl_rad.astype("float64"), PX, np.float64(n_top))  # example.py:69
    r["ti"], r["mti"] = ti.astype("float32"), mti.astype("float32")
    r["downslope"] = D.downsloper(dem, fdr, PX, delta)  # example.py:74
    fdist, idx, hand = F.flow_hand_index(dem, fdr, river, PX)  # example.py:82
    r["fdist"], r["idx"], r["hand"] = fdist, idx.astype("int64"), hand
    hand_in = hand.astype("float64") if hand.dtype == np.float32 else hand
    fac_c, idx_c = np.ascontiguousarray(fac), np.ascontiguousarray(idx.astype("int64"))
    r["gfi"] = G.gfi_calculator(hand_in, fac_c, idx_c, np.float64(n_gfi), np.float64(b), PX).astype("float32")
    r["lnhlh"] = G.ln_hl_H_calculator(hand_in, fac_c, np.float64(n_gfi), np.float64(b), PX).astype("float32")
    return r


def make_cudasim_example_crop():
    dem, fdr, fac, _, _ = load_example()
    # window with rivers, hillslopes and a nodata corner (bottom-left of the basin clip)
    r0, c0, h, w = 1000, 600, 96, 112
    best = None
    for rr in range(300, 1900, 100):
        for cc in range(100, 1400, 100):
            f = fac[rr:rr + h, cc:cc + w]
            nd = (fdr[rr:rr + h, cc:cc + w] == 0).mean()
            score = (f > 2000).sum() * (0.02 < nd < 0.4)
            if best is None or score > best[0]:
                best = (score, rr, cc)
    _, r0, c0 = best
    dem, fdr, fac = dem[r0:r0 + h, c0:c0 + w].copy(), fdr[r0:r0 + h, c0:c0 + w].copy(), fac[r0:r0 + h, c0:c0 + w].copy()
    river = np.where(fac > 2000, 1, 0).astype("int8")
    print("crop", r0, c0, "river cells", river.sum(), "nodata", (fdr == 0).sum())
    res = run_reference_gpu_entry_points(dem, fdr, fac, river, 0.1, 5, 0.4, 0.1)
    np.savez_compressed(os.path.join(HERE, "cudasim_example_crop.npz"), dem=dem, fdr=fdr, fac=fac, river=river,
                        origin=np.array([r0, c0]), **res)


def make_cudasim_synth_f32():
    rows, cols = 72, 88
    dem = oracle.conditioned_dem(rows, cols, seed=7)
    rng = np.random.default_rng(5)
    # nodata holes + a nodata edge strip; one value below -100 (centre test is <=, slope.py:231)
    dem[10:16, 20:27] = -100
    dem[:, :3] = -100
    dem[40, 60] = -100
    dem[55, 30] = -250.0
    dem += 0  # keep f32
    _, d8 = oracle.slope_d8(dem, float(PX))
    acc, left = oracle.flow_accumulation(d8)
    assert left == 0
    river = np.where(acc > 60, 1, 0).astype("int8")
    print("synth river cells", river.sum())
    res = run_reference_gpu_entry_points(dem, d8, acc, river, 0.1, 0.5, 0.4, 0.1)
    np.savez_compressed(os.path.join(HERE, "cudasim_synth_f32.npz"), dem=dem, fdr=d8, fac=acc, river=river, **res)
    del rng


def make_cudasim_special():
    """Hand-made D8 grids: cycles, unknown codes, code 0 on the path, river cell with code 0,
    walks leaving through every edge and corner (flowhand.py:623-837; downslope.py:468-529)."""
    import descriptools.flowhand as F
    import descriptools.downslope as D

    rows, cols = 12, 14
    fdr = np.full((rows, cols), 1, np.uint8)  # everything flows east ...
    fdr[:, -1] = 4  # ... then south along the last column
    river = np.zeros((rows, cols), np.int8)
    river[rows - 1, cols - 1] = 1
    dem = (200 - 3 * np.arange(cols)[None, :] - 2 * np.arange(rows)[:, None]).astype(np.float32)
    # 2-cycle, 3-cycle, 4-cycle
    fdr[1, 1], fdr[1, 2] = 1, 16
    fdr[3, 1], fdr[3, 2], fdr[4, 1] = 1, 8, 64
    fdr[6, 1], fdr[6, 2], fdr[7, 2], fdr[7, 1] = 1, 4, 16, 64
    fdr[1, 0] = 1  # feeds the 2-cycle
    # unknown codes, code 0 in the middle of a row, river cell carrying code 0
    fdr[2, 5] = 3
    fdr[5, 6] = 0
    fdr[8, 7] = 0
    river[8, 7] = 1
    fdr[9, 4] = 255
    # exits through every edge / corner
    fdr[0, 3], fdr[0, 5], fdr[0, 7] = 32, 64, 128
    fdr[0, 0], fdr[0, cols - 1] = 32, 128
    fdr[rows - 1, 0], fdr[rows - 1, 2], fdr[rows - 1, 4], fdr[rows - 1, 6] = 8, 2, 4, 8
    fdr[4, 0], fdr[5, 0], fdr[6, 0] = 16, 32, 8
    fdr[2, cols - 1], fdr[3, cols - 1], fdr[4, cols - 1] = 1, 2, 128
    # a second river reached diagonally
    river[10, 9] = 1
    fdr[9, 8] = 2
    dem[5, 6] = -100
    dem[8, 7] = -100
    fdist, idx, hand = F.flow_hand_index(dem, fdr, river, PX)
    # a valid-DEM cell that never moves makes the reference raise ZeroDivisionError
    # (downslope.py:312); mask the two unknown-code cells for the downslope run
    dem_ds = dem.copy()
    dem_ds[2, 5] = dem_ds[9, 4] = -100
    down = D.downsloper(dem_ds, fdr, PX, 4)
    np.savez_compressed(os.path.join(HERE, "cudasim_special.npz"), dem=dem, dem_ds=dem_ds, fdr=fdr, river=river,
                        fdist=fdist, idx=idx.astype("int64"), hand=hand, downslope=down)
    print("special: failed cells", int((idx == -100).sum()), "of", idx.size)


def make_cudasim_cap():
    """Serpentine D8 path longer than the 20 000-move cap (flowhand.py:835): only the first
    `threads` cells are launched (blocks=1) so the simulator finishes in seconds."""
    import descriptools.flowhand as F

    rows, cols = 52, 400
    fdr = np.zeros((rows, cols), np.uint8)
    for r in range(rows):
        fdr[r, :] = 1 if r % 2 == 0 else 16
        fdr[r, -1 if r % 2 == 0 else 0] = 4
    path = []
    r, c = 0, 0
    while True:
        path.append((r, c))
        f = fdr[r, c]
        if f == 1: c += 1
        elif f == 16: c -= 1
        else: r += 1
        if r >= rows: break
    river = np.zeros((rows, cols), np.int8)
    rr, cc = path[20001]  # cell 0 needs 20001 moves (fails), cell 1 needs 20000 (succeeds)
    river[rr, cc] = 1
    dem = np.zeros((rows, cols), np.float32) + 50
    threads = 6
    fdist, idx = F.flow_distance_index_cpu(dem, fdr, river, PX, np.zeros((4, 1)), np.zeros((4, 1)), np.zeros(4),
                                           0, 0, cols, blocks=1, threads=threads)
    np.savez_compressed(os.path.join(HERE, "cudasim_cap.npz"), fdr=fdr, river=river, n_cells=np.array(threads),
                        fdist=fdist.reshape(-1)[:threads].astype("float32"),
                        idx=idx.reshape(-1)[:threads].astype("int64"))
    print("cap:", idx.reshape(-1)[:threads], fdist.reshape(-1)[:threads])


STEPS = {
    "inputs": make_example_inputs,
    "cpujit": make_example_cpujit,
    "crop": make_cudasim_example_crop,
    "synth": make_cudasim_synth_f32,
    "special": make_cudasim_special,
    "cap": make_cudasim_cap,
}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*", default=list(STEPS))
    for name in ap.parse_args().only:
        print("==", name)
        STEPS[name]()
