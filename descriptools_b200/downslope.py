"""downslope.py -- drop-in for descriptools/downslope.py (reference: downslope.py:317-532)."""
import numpy as np

from . import device
from ._convert import dem_to_native, fdr_to_u8, to_dev, to_host


def downsloper(dem, flow_direction, px, elevation_difference, column_division=0, row_division=0):
    """Downslope index (float32, not in percent) -- downslope.py:317-376.

    The reference runs a GPU pass and then a single-threaded CPU pass over the cells the kernel
    flagged -50 (downslope.py:373-374); one kernel produces that composite here.
    """
    out = device.downslope(to_dev(dem_to_native(dem)), to_dev(fdr_to_u8(flow_direction)), px, elevation_difference)
    return to_host(out)


def downslope_cpu(dem, flow_direction, px, elevation_difference, blocks=0, threads=0):
    """Host wrapper -- downslope.py:379-431.  Returns float64 like the reference (downslope.py:418).

    Deviation: the reference's kernel leaves -50 in cells it cannot finish and relies on the caller's
    CPU pass; this wrapper returns the finished (composite) values, so no -50 ever appears.
    """
    return downsloper(dem, flow_direction, px, elevation_difference).astype(np.float64)
