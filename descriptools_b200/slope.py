"""slope.py -- drop-in for descriptools/slope.py (reference: slope.py:96-259).

Same function names, positional order and return dtypes as the reference; the work is done by
the fused slope+D8 stencil kernel (csrc/slope_d8.cu) through libdtb200's C ABI.
"""
import numpy as np

from . import device
from ._convert import dem_to_native, to_dev, to_host


def sloper(dem, px, division_column=0, division_row=0):
    """Highest slope (percent) to a neighbouring cell -- slope.py:96-149.

    Returns a float64 array holding float32 values, like the reference (slope.py:119,145).
    `division_*` are accepted and ignored: the result equals the unpartitioned reference path.
    """
    d = dem_to_native(dem)
    slope, _ = device.slope_d8(to_dev(d), px, want_slope=True, want_d8=False)
    return to_host(slope).astype(np.float64)


def slope_cpu(dem, px, extra, blocks=0, threads=0):
    """Host wrapper of the slope kernel -- slope.py:152-206.

    `extra[k] == 1` (up, left, right, down) pads that side with -100; the outer ring of the
    padded tile is then discarded.  Returns float32 like the reference (slope.py:193,200-205).
    `blocks` / `threads` are accepted for signature parity and ignored.
    """
    d = dem_to_native(dem)
    nd = np.asarray(-100, dtype=d.dtype)
    pad = ((1 if extra[0] == 1 else 0, 1 if extra[3] == 1 else 0), (1 if extra[1] == 1 else 0, 1 if extra[2] == 1 else 0))
    d = np.ascontiguousarray(np.pad(d, pad, constant_values=nd))  # slope.py:175-182
    slope, _ = device.slope_d8(to_dev(d), px, want_slope=True, want_d8=False)
    return to_host(slope)[1:-1, 1:-1].copy()  # slope.py:202-205
