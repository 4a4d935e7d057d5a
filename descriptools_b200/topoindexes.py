"""topoindexes.py -- drop-in for descriptools/topoindexes.py (reference: topoindexes.py:109-295)."""
import numpy as np

from . import device
from ._convert import _exact_cast, ints_to_native, to_dev, to_host


def _slope_f32(slope):
    s = np.asarray(slope)
    if s.dtype != np.float32:
        s = _exact_cast(s, np.float32, "slope")  # the example passes float32 radians (example.py:63-64)
    return np.ascontiguousarray(s)


def topographic_index(flow_accumulation, slope, px, n_top, div_col=0, div_row=0):
    """Topographic index and modified topographic index -- topoindexes.py:109-167.

    `slope` is in radians (example.py:63).  Returns two float64 arrays of float32 values
    (topoindexes.py:146-147).  Formula of the reference's GPU kernels: tan(beta + 0.01).
    """
    ti, mti = topographic_index_cpu(flow_accumulation, slope, px, n_top)
    return ti.astype(np.float64), mti.astype(np.float64)


def topographic_index_cpu(flow_accumulation, slope, px, expoent, blocks=0, threads=0):
    """Host wrapper of both kernels -- topoindexes.py:170-230 (float32, float32)."""
    fac = to_dev(ints_to_native(flow_accumulation, "flow_accumulation"))
    ti, mti = device.ti_mti(fac, to_dev(_slope_f32(slope)), px, expoent)
    return to_host(ti), to_host(mti)
