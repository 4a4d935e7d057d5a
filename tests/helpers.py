"""Shared test helpers: golden loaders, the example.py glue and a NumPy restatement of the
reference's evaluation.py steps used by KAT-1 (evaluation.py:5-9, 12-87, 90-123, 126-171)."""
import hashlib
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def example_inputs():
    z = load("example_inputs.npz")
    shape = tuple(int(v) for v in z["shape"])
    flood = np.unpackbits(z["flood_bits"])[: shape[0] * shape[1]].reshape(shape).astype(np.int8)
    fac = z["fac"].astype(np.int64)
    river = np.where(fac > 128000, 1, 0).astype(np.int8)  # example.py:52
    return dict(dem=z["dem"], fdr=z["fdr"], fac=fac, river=river, flood=flood, hand_class=z["hand_class"])


# ---- evaluation.py restatement (KAT-1 only) -------------------------------------------
def min_max_scale(mat, mn, mx, nodata):  # evaluation.py:5-9
    scaled = np.where(mat == nodata, np.nan, mat)
    return np.where(np.isnan(mat), scaled, (scaled - mn) / (mx - mn))


def binary_map(desc, th, under=True):  # evaluation.py:90-123
    d = np.where(desc == desc[0, 0], np.nan, desc)
    hit = (d <= th) if under else (d >= th)
    return np.where(np.isnan(d), 0, np.where(hit, 1, 0))


def avaliacao(binary, flood):  # evaluation.py:126-171
    cmp_ = np.where(flood == 1, 2, flood)
    cls = binary + cmp_
    tn, fp, fn, tp = [(cls == k).sum() for k in range(4)]
    c = tp / (tp + fn)
    f = tp / (tp + fn + fp)
    return c, f, cls


def calibration(desc, flood):  # evaluation.py:12-87, direction 'under'
    fit = lambda t: avaliacao(binary_map(desc, t), flood)[1]
    f1, f2, f3 = fit(0.25), fit(0.50), fit(0.75)
    if f3 > f2:
        it, best = (75, f3) if f3 > f1 else (25, f1)
    else:
        it, best = (50, f2) if f2 > f1 else (25, f1)
    th = None
    for i in range(it - 20, it + 30, 10):
        v = fit(i / 100)
        if v >= best:
            best, th = v, i
    it = th
    for i in range(it - 5, it + 6):
        v = fit(i / 100)
        if v > best:
            best, th = v, i
    for scale in (1000, 10000):
        it = th * 10
        th = it
        for i in range(it - 10, it + 11):
            v = fit(i / scale)
            if v > best:
                best, th = v, i
    return th / 10000
