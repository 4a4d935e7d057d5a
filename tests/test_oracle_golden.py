"""Pin the CPU oracle (oracle/) against vectors produced by the reference itself.

CPU-only.  Sources of truth (tests/golden/make_golden.py):
  * example_cpujit.npz  -- the reference's compiled CPU-jit twins on the full bundled example
  * cudasim_*.npz       -- the reference's unmodified @cuda.jit kernels under CUDASIM
  * example_inputs.npz  -- bundled rasters incl. KAT-1 (hand_class.tif) and KAT-2 (fdr -> fac)
"""
import numpy as np
import pytest

import oracle
from helpers import avaliacao, binary_map, calibration, example_inputs, load, min_max_scale, sha

PX = 12.5
RTOL, ATOL = 1e-5, 1e-6  # north-star float tolerance (SURVEY.md addendum: add atol for values crossing 0)


@pytest.fixture(scope="module")
def ex():
    return example_inputs()


@pytest.fixture(scope="module")
def ex_hand(ex):
    return oracle.flow_hand_index(ex["dem"], ex["fdr"], ex["river"], PX)


def _check_sha(gold, name, arr):
    assert str(arr.dtype) == str(gold[name + "_dtype"])
    np.testing.assert_array_equal(arr.reshape(-1)[::37], gold[name + "_sample"])
    assert sha(arr) == str(gold[name + "_sha256"])


def test_full_example_slope_bit_exact(ex):
    gold = load("example_cpujit.npz")
    slope, _ = oracle.slope_d8(ex["dem"], PX)
    _check_sha(gold, "slope", slope)


def test_full_example_hand_bit_exact(ex, ex_hand):
    gold = load("example_cpujit.npz")
    fdist, idx, hand = ex_hand
    _check_sha(gold, "idx", idx)
    _check_sha(gold, "fdist", fdist)
    _check_sha(gold, "hand", hand)


def test_full_example_downslope_bit_exact(ex):
    gold = load("example_cpujit.npz")
    _check_sha(gold, "downslope", oracle.downslope(ex["dem"], ex["fdr"], PX, 5))


def test_full_example_gfi_lnhlh(ex, ex_hand):
    gold = load("example_cpujit.npz")
    _, idx, hand = ex_hand
    _check_sha(gold, "gfi", oracle.gfi(hand, ex["fac"], idx, 0.4, 0.1, PX))
    _check_sha(gold, "lnhlh", oracle.ln_hl_H(hand, ex["fac"], 0.4, 0.1, PX))


def test_kat1_hand_class(ex, ex_hand):
    """example.py:106-147 -> Example/output/hand_class.tif, the reference's only golden output."""
    _, _, hand = ex_hand
    elements = np.unique(hand)
    mn, mx = elements[1], elements[-1]  # example.py:113-115
    assert (mn, mx) == (0, 259)
    desc = min_max_scale(hand, mn, mx, -100)
    th = calibration(desc, ex["flood"])
    assert th == pytest.approx(0.012)
    c, f, cls = avaliacao(binary_map(desc, th), ex["flood"])
    assert c == pytest.approx(0.8581615676712259, abs=1e-12)
    assert f == pytest.approx(0.7240945135019289, abs=1e-12)
    np.testing.assert_array_equal(cls.astype(np.uint8), ex["hand_class"])


def test_kat2_flow_accumulation_convention(ex):
    """12_fdr.tif -> 12_fac.tif: fac counts strictly-upstream cells (self excluded); the
    fixture is a clip, so cells downstream of the clip edge carry extra external inflow."""
    fdr, fac = ex["fdr"], ex["fac"]
    acc, left = oracle.flow_accumulation(fdr)
    assert left == 0
    valid = fdr != 0
    assert (acc[~valid] == -100).all() and (fac[~valid] == -100).all()
    assert (acc[valid] <= fac[valid]).all()
    assert (acc[valid] == fac[valid]).mean() > 0.98
    # external inflow e(p) = fac[p] - sum_{q->p} (fac[q] + 1) is >= 0 everywhere and non-zero only
    # next to nodata / the raster edge; accumulating it down the tree reproduces fac exactly.
    rows, cols = fdr.shape
    off = {1: (0, 1), 2: (1, 1), 4: (1, 0), 8: (1, -1), 16: (0, -1), 32: (-1, -1), 64: (-1, 0), 128: (-1, 1)}
    rr, cc = np.nonzero(valid)
    dr = np.zeros(rr.size, np.int64)
    dc = np.zeros(rr.size, np.int64)
    for code, (a, b) in off.items():
        m = fdr[rr, cc] == code
        dr[m], dc[m] = a, b
    tr, tc = rr + dr, cc + dc
    ok = (tr >= 0) & (tr < rows) & (tc >= 0) & (tc < cols)
    ok[ok] &= valid[tr[ok], tc[ok]]
    inflow = np.zeros((rows, cols), np.int64)
    np.add.at(inflow, (tr[ok], tc[ok]), fac[rr[ok], cc[ok]] + 1)
    ext = np.where(valid, fac - inflow, 0)
    assert (ext >= 0).all()
    pad = np.pad(~valid, 1, constant_values=True)
    near_nd = np.zeros_like(valid)
    for a in (0, 1, 2):
        for b in (0, 1, 2):
            near_nd |= pad[a:a + rows, b:b + cols]
    assert not (ext[~near_nd] != 0).any()
    # accumulate ext down the tree: total = acc + (sum of ext over the upstream closure incl. self)
    order = np.argsort(acc[valid], kind="stable")  # upstream cells have strictly smaller acc
    tot = ext.astype(np.int64).copy()
    nxt = np.full(rows * cols, -1, np.int64)
    nxt[(rr[ok] * cols + cc[ok])] = tr[ok] * cols + tc[ok]
    flat_tot = tot.reshape(-1)
    lin = (rr * cols + cc)[order]
    for p in lin:
        q = nxt[p]
        if q >= 0:
            flat_tot[q] += flat_tot[p]
    np.testing.assert_array_equal((acc + tot)[valid], fac[valid])


@pytest.mark.parametrize("name", ["cudasim_example_crop.npz", "cudasim_synth_f32.npz"])
def test_cudasim_entry_points(name):
    """Oracle == the reference's GPU kernels (CUDASIM) through its public entry points."""
    g = load(name)
    dem, fdr, fac, river = g["dem"], g["fdr"], g["fac"], g["river"]
    delta = 5 if dem.dtype == np.int16 else 0.5
    slope, _ = oracle.slope_d8(dem, PX)
    np.testing.assert_array_equal(slope, g["slope"])
    fdist, idx, hand = oracle.flow_hand_index(dem, fdr, river, PX)
    np.testing.assert_array_equal(idx, g["idx"])
    np.testing.assert_array_equal(fdist, g["fdist"])
    np.testing.assert_array_equal(hand, g["hand"].astype(hand.dtype))
    assert hand.dtype == g["hand"].dtype
    np.testing.assert_array_equal(oracle.downslope(dem, fdr, PX, delta), g["downslope"])
    ti, mti = oracle.ti_mti(fac, g["slope_rad"], PX, 0.1)
    np.testing.assert_allclose(ti, g["ti"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(mti, g["mti"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(oracle.gfi(hand, fac, idx, 0.4, 0.1, PX), g["gfi"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(oracle.ln_hl_H(hand, fac, 0.4, 0.1, PX), g["lnhlh"], rtol=RTOL, atol=ATOL)
    # the pointwise kernels round-trip through libm on both sides: expect (near) bit equality
    for a, b in ((ti, g["ti"]), (mti, g["mti"])):
        assert (a != b).mean() < 1e-3


def test_cudasim_special_cases():
    """cycles, unknown codes, code 0 on the path, river with code 0, exits via every edge."""
    g = load("cudasim_special.npz")
    fdist, idx, hand = oracle.flow_hand_index(g["dem"], g["fdr"], g["river"], PX)
    np.testing.assert_array_equal(idx, g["idx"])
    np.testing.assert_array_equal(fdist, g["fdist"])
    np.testing.assert_array_equal(hand, g["hand"])
    np.testing.assert_array_equal(oracle.downslope(g["dem_ds"], g["fdr"], PX, 4), g["downslope"])


def test_cudasim_move_cap():
    """flowhand.py:835: a walk needing 20001 moves fails, 20000 succeeds."""
    g = load("cudasim_cap.npz")
    n = int(g["n_cells"])
    fdist, idx = oracle.flow_distance_index(g["fdr"], g["river"], PX)
    np.testing.assert_array_equal(idx.reshape(-1)[:n], g["idx"])
    np.testing.assert_array_equal(fdist.reshape(-1)[:n], g["fdist"])
    assert g["idx"][0] == -100 and g["idx"][1] != -100


def test_d8_spec_properties():
    """A2 (frozen here): every valid cell of a conditioned DEM gets a non-zero code, interior
    cells point at a strictly lower neighbour, and the direction attains the slope maximum."""
    dem = oracle.conditioned_dem(96, 130, seed=3)
    dem[20:30, 40:55] = -100
    dem = oracle.priority_flood_eps(dem)
    slope, d8 = oracle.slope_d8(dem, PX)
    valid = dem > -100
    assert (d8[valid] != 0).all() and (d8[~valid] == 0).all()
    acc, left = oracle.flow_accumulation(d8)
    assert left == 0
    off = {1: (0, 1), 2: (1, 1), 4: (1, 0), 8: (1, -1), 16: (0, -1), 32: (-1, -1), 64: (-1, 0), 128: (-1, 1)}
    rows, cols = dem.shape
    for r in range(1, rows - 1):
        for c in range(1, cols - 1):
            if not valid[r, c]:
                continue
            dr, dc = off[int(d8[r, c])]
            zq = dem[r + dr, c + dc]
            if zq == -100:
                assert slope[r, c] == 0
                continue
            assert zq < dem[r, c]
            g = (np.float32(dem[r, c] - zq)).astype(np.float64) / (PX if 0 in (dr, dc) else PX * np.sqrt(2.0))
            assert np.float32(g * 100.0) == slope[r, c]


def test_band_rows_equal_full():
    dem = oracle.conditioned_dem(64, 48, seed=11)
    s, d = oracle.slope_d8(dem, PX)
    s2, d2 = oracle.slope_d8(dem, PX, 10, 37)
    np.testing.assert_array_equal(s[10:37], s2)
    np.testing.assert_array_equal(d[10:37], d2)
