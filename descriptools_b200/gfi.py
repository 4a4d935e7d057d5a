"""gfi.py -- drop-in for descriptools/gfi.py (reference: gfi.py:118-440)."""
import numpy as np

from . import device
from ._convert import dem_to_native, ints_to_native, to_dev, to_host


def river_accumulation(flow_accumulation, indices):
    """Flow accumulation of each cell's river cell -- gfi.py:118-147 (dtype of flow_accumulation)."""
    fac_in = np.asarray(flow_accumulation)
    fac = ints_to_native(fac_in, "flow_accumulation")
    idx = ints_to_native(np.asarray(indices).reshape(fac.shape), "indices")
    out = to_host(device.river_accumulation(to_dev(fac), to_dev(idx)))
    return out.astype(fac_in.dtype) if out.dtype != fac_in.dtype else out


def gfi_calculator(hand, flow_accumulation, indices, n_gfi, scale_factor, size, division_column=0, division_row=0):
    """Geomorphic flood index -- gfi.py:150-207.  float64 array of float32 values (gfi.py:188)."""
    h = to_dev(dem_to_native(hand))
    fac = to_dev(ints_to_native(flow_accumulation, "flow_accumulation"))
    idx = to_dev(ints_to_native(np.asarray(indices).reshape(tuple(h.shape)), "indices"))
    racc = device.river_accumulation(fac, idx)  # gfi.py:186
    return to_host(device.gfi(h, racc, n_gfi, scale_factor, size)).astype(np.float64)


def geomorphic_flood_index_cpu(hand, river_flow_accumulation, expoent, scale_factor, size, blocks=0, threads=0):
    """Host wrapper of the GFI kernel -- gfi.py:210-264 (float32)."""
    h = to_dev(dem_to_native(hand))
    racc = to_dev(ints_to_native(river_flow_accumulation, "river_flow_accumulation"))
    return to_host(device.gfi(h, racc, expoent, scale_factor, size))


def ln_hl_H_calculator(hand, flow_accumulation, n_gfi, scale_factor, size, division_column=0, division_row=0):
    """ln(hl/H) -- gfi.py:297-346.  float64 array of float32 values (gfi.py:329)."""
    return ln_hl_H_cpu(hand, flow_accumulation, n_gfi, scale_factor, size).astype(np.float64)


def ln_hl_H_cpu(hand, flow_accumulation, expoent, scale_factor, size, blocks=0, threads=0):
    """Host wrapper of the ln(hl/H) kernel -- gfi.py:349-400 (float32)."""
    h = to_dev(dem_to_native(hand))
    fac = to_dev(ints_to_native(flow_accumulation, "flow_accumulation"))
    return to_host(device.ln_hl_H(h, fac, expoent, scale_factor, size))
