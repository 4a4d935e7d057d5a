"""CPU oracle for the descriptools terrain-descriptor path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker or the
reported CPU baseline.  ``descriptools_b200`` never imports it.

The arithmetic lives in ``dt_oracle.c`` / ``dt_condition.cpp`` (each function cites
the reference file:line it restates); this module is a thin ctypes + NumPy binding
that mirrors the reference's array conventions (2-D row-major arrays, -100 nodata).
Parity pin: ``tests/test_oracle_golden.py`` (vectors made by the reference itself,
``tests/golden/make_golden.py``).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_LIBS = {}

MAX_MOVES_HAND = 20000  # flowhand.py:835
MAX_MOVES_DOWNSLOPE = 5000  # downslope.py:303


def build(force: bool = False) -> None:
    """Compile the oracle libraries with gcc/g++ (Makefile in this directory)."""
    want = [os.path.join(_BUILD, "libdt_oracle.so"), os.path.join(_BUILD, "libdt_condition.so")]
    srcs = [os.path.join(_HERE, "dt_oracle.c"), os.path.join(_HERE, "dt_condition.cpp")]
    fresh = all(
        os.path.exists(w) and os.path.getmtime(w) >= os.path.getmtime(s) for w, s in zip(want, srcs)
    )
    if fresh and not force:
        return
    subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))


def _lib(name: str) -> ctypes.CDLL:
    if name not in _LIBS:
        path = os.path.join(_BUILD, name)
        if not os.path.exists(path):
            build()
        _LIBS[name] = ctypes.CDLL(path)
    return _LIBS[name]


def set_threads(n: int) -> int:
    """OpenMP team size of the oracle's loops; returns the value in effect."""
    lib = _lib("libdt_oracle.so")
    lib.orc_set_threads(ctypes.c_int(int(n)))
    return int(lib.orc_max_threads())


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dtype)


_i64 = ctypes.c_int64
_f64 = ctypes.c_double


def slope_d8(dem: np.ndarray, px: float, row_begin: int = 0, row_end: int | None = None):
    """(slope f32 in %, d8 u8) for rows [row_begin,row_end).  slope.py:228-259 + SURVEY A2."""
    lib = _lib("libdt_oracle.so")
    rows, cols = dem.shape
    row_end = rows if row_end is None else row_end
    if dem.dtype == np.int16:
        fn, d = lib.orc_slope_d8_i16, _c(dem, np.int16)
    else:
        fn, d = lib.orc_slope_d8_f32, _c(dem, np.float32)
    slope = np.empty((row_end - row_begin, cols), np.float32)
    d8 = np.empty((row_end - row_begin, cols), np.uint8)
    fn(_p(d), _i64(rows), _i64(cols), _i64(row_begin), _i64(row_end), _f64(px), _p(slope), _p(d8))
    return slope, d8


def flow_accumulation(d8: np.ndarray, nodata_fill: int = -100):
    """(acc int64, n_unfinalised).  SURVEY App. A3."""
    lib = _lib("libdt_oracle.so")
    lib.orc_flowacc.restype = _i64
    d = _c(d8, np.uint8)
    rows, cols = d.shape
    acc = np.empty((rows, cols), np.int64)
    left = lib.orc_flowacc(_p(d), _i64(rows), _i64(cols), _p(acc), _i64(nodata_fill))
    return acc, int(left)


def flow_distance_index(fdr: np.ndarray, river: np.ndarray, px: float, max_moves: int = MAX_MOVES_HAND):
    """(flow_distance f32, indices int64).  flowhand.py:599-846 (unpartitioned)."""
    lib = _lib("libdt_oracle.so")
    f, rv = _c(fdr, np.uint8), _c(river, np.int8)
    rows, cols = f.shape
    fdist = np.empty((rows, cols), np.float32)
    idx = np.empty((rows, cols), np.int64)
    lib.orc_flow_distance_index(_p(f), _p(rv), _i64(rows), _i64(cols), _f64(px), _i64(max_moves), _p(fdist), _p(idx))
    return fdist, idx


def hand(dem: np.ndarray, idx: np.ndarray):
    """HAND in the DEM's dtype.  flowhand.py:431-442."""
    lib = _lib("libdt_oracle.so")
    ix = _c(idx, np.int64)
    if dem.dtype == np.int16:
        d, fn, out = _c(dem, np.int16), lib.orc_hand_i16, np.empty(dem.shape, np.int16)
    else:
        d, fn, out = _c(dem, np.float32), lib.orc_hand_f32, np.empty(dem.shape, np.float32)
    fn(_p(d), _p(ix), _i64(d.size), _p(out))
    return out


def flow_hand_index(dem, fdr, river, px, max_moves: int = MAX_MOVES_HAND):
    """flowhand.py:242-411 with division_* = 0."""
    fdist, idx = flow_distance_index(fdr, river, px, max_moves)
    return fdist, idx, hand(dem, idx)


def downslope(dem, fdr, px, delta, max_moves: int = MAX_MOVES_DOWNSLOPE):
    """Composite downsloper (GPU pass + CPU fix-up).  downslope.py:458-532 + 194-312."""
    lib = _lib("libdt_oracle.so")
    f = _c(fdr, np.uint8)
    rows, cols = f.shape
    out = np.empty((rows, cols), np.float32)
    if dem.dtype == np.int16:
        d, fn = _c(dem, np.int16), lib.orc_downslope_i16
    else:
        d, fn = _c(dem, np.float32), lib.orc_downslope_f32
    fn(_p(d), _p(f), _i64(rows), _i64(cols), _f64(px), _f64(delta), _i64(max_moves), _p(out))
    return out


def river_accumulation(fac, idx):
    """gfi.py:136-147."""
    lib = _lib("libdt_oracle.so")
    a, ix = _c(fac, np.int64), _c(idx, np.int64)
    out = np.empty(a.shape, np.int64)
    lib.orc_river_accumulation(_p(a), _p(ix), _i64(a.size), _p(out))
    return out


def _hand_arg(hand_arr, lib, stem):
    if hand_arr.dtype == np.int16:
        return _c(hand_arr, np.int16), getattr(lib, stem + "_i16")
    return _c(hand_arr, np.float32), getattr(lib, stem + "_f32")


def gfi(hand_arr, fac, idx, n, b, size):
    """gfi_calculator: gather (gfi.py:136-147) + kernel (gfi.py:287-294).  f32."""
    lib = _lib("libdt_oracle.so")
    racc = river_accumulation(fac, idx)
    h, fn = _hand_arg(hand_arr, lib, "orc_gfi")
    out = np.empty(h.shape, np.float32)
    fn(_p(h), _p(racc), _i64(h.size), _f64(n), _f64(b), _f64(size), _p(out))
    return out


def ln_hl_H(hand_arr, fac, n, b, size):
    """gfi.py:427-440.  f32."""
    lib = _lib("libdt_oracle.so")
    a = _c(fac, np.int64)
    h, fn = _hand_arg(hand_arr, lib, "orc_lnhlh")
    out = np.empty(h.shape, np.float32)
    fn(_p(h), _p(a), _i64(h.size), _f64(n), _f64(b), _f64(size), _p(out))
    return out


def ti_mti(fac, slope_rad, px, n):
    """topoindexes.py:250-261, 284-295 (the *_gpu formula).  (TI f32, MTI f32)."""
    lib = _lib("libdt_oracle.so")
    a, s = _c(fac, np.int64), _c(slope_rad, np.float32)
    ti = np.empty(a.shape, np.float32)
    mti = np.empty(a.shape, np.float32)
    lib.orc_ti_mti(_p(a), _p(s), _i64(a.size), _f64(px), _f64(n), _p(ti), _p(mti))
    return ti, mti


# ---- synthetic DEM recipe "dtb-synth-v1" (dt_condition.cpp) ----
SYNTH_SEED = 20260101
SYNTH_Z0, SYNTH_SR, SYNTH_SC, SYNTH_DEPTH = 100.0, 0.02, 0.005, 6.0
SYNTH_AMP0, SYNTH_HURST = 15.0, 0.6


def synth_amplitudes() -> np.ndarray:
    k = np.arange(10, dtype=np.float64)
    return (SYNTH_AMP0 * (0.5 ** k) ** SYNTH_HURST).astype(np.float32)


def synth_dem(rows: int, cols: int, row0: int = 0, seed: int = SYNTH_SEED) -> np.ndarray:
    lib = _lib("libdt_condition.so")
    out = np.empty((rows, cols), np.float32)
    amp = synth_amplitudes()
    f32 = ctypes.c_float
    lib.orc_synth_dem_f32(_i64(rows), _i64(cols), _i64(row0), ctypes.c_uint32(seed), _p(amp),
                          f32(SYNTH_Z0), f32(SYNTH_SR), f32(SYNTH_SC), f32(SYNTH_DEPTH), _p(out))
    return out


def priority_flood_eps(dem: np.ndarray) -> np.ndarray:
    lib = _lib("libdt_condition.so")
    out = np.array(dem, dtype=np.float32, order="C", copy=True)
    lib.orc_priority_flood_eps_f32(_p(out), _i64(out.shape[0]), _i64(out.shape[1]))
    return out


def conditioned_dem(rows: int, cols: int, seed: int = SYNTH_SEED) -> np.ndarray:
    """Synthetic hydrologically conditioned DEM (host path: generate + priority-flood)."""
    return priority_flood_eps(synth_dem(rows, cols, 0, seed))
