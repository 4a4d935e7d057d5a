"""Host <-> device marshalling for the NumPy drop-in modules: dtype normalisation that never
changes a value (anything that cannot be represented exactly raises instead of rounding)."""
from __future__ import annotations

import numpy as np
import torch

from .device import require_cuda


def _exact_cast(a: np.ndarray, dtype, what: str) -> np.ndarray:
    b = a.astype(dtype)
    with np.errstate(invalid="ignore"):
        same = (b == a) | ((a != a) & (b != b)) if np.issubdtype(a.dtype, np.floating) else (b == a)
    if not np.all(same):
        raise TypeError(f"{what}: values of dtype {a.dtype} are not exactly representable as {np.dtype(dtype)}; "
                        "descriptools_b200 computes on int16 or float32 elevations")
    return b


def dem_to_native(dem) -> np.ndarray:
    """int16 stays int16, float32 stays float32; other dtypes are converted only if exact."""
    dem = np.asarray(dem)
    if dem.ndim != 2:
        raise ValueError("raster must be 2-D")
    if dem.dtype == np.int16 or dem.dtype == np.float32:
        return np.ascontiguousarray(dem)
    if np.issubdtype(dem.dtype, np.integer) or dem.dtype == np.bool_:
        if dem.size and dem.min() >= -32768 and dem.max() <= 32767:
            return np.ascontiguousarray(dem.astype(np.int16))
        return np.ascontiguousarray(_exact_cast(dem, np.float32, "dem"))
    if np.issubdtype(dem.dtype, np.floating):
        return np.ascontiguousarray(_exact_cast(dem, np.float32, "dem"))
    raise TypeError(f"unsupported raster dtype {dem.dtype}")


def fdr_to_u8(fdr) -> np.ndarray:
    fdr = np.asarray(fdr)
    if fdr.dtype == np.uint8:
        return np.ascontiguousarray(fdr)
    if fdr.dtype == np.int8:
        return np.ascontiguousarray(fdr).view(np.uint8)
    if fdr.size and (fdr.min() < 0 or fdr.max() > 255):
        # the reference treats codes <= 0 as nodata (flowhand.py:601) and ignores unknown codes
        fdr = np.where((fdr < 0) | (fdr > 255), 0, fdr)
    return np.ascontiguousarray(_exact_cast(np.asarray(fdr), np.uint8, "flow_direction"))


def river_to_i8(river) -> np.ndarray:
    river = np.asarray(river)
    if river.dtype == np.int8:
        return np.ascontiguousarray(river)
    return np.ascontiguousarray((river == 1).astype(np.int8))  # only `== 1` is ever tested (flowhand.py:609,622)


def ints_to_native(a, what: str) -> np.ndarray:
    """accumulation / index rasters: int32 when it fits, else int64."""
    a = np.asarray(a)
    if np.issubdtype(a.dtype, np.floating):
        a = _exact_cast(a, np.int64, what)
    elif not np.issubdtype(a.dtype, np.integer):
        raise TypeError(f"{what}: unsupported dtype {a.dtype}")
    if a.dtype == np.int32:
        return np.ascontiguousarray(a)
    if a.size and a.min() >= -(2**31) and a.max() < 2**31:
        return np.ascontiguousarray(a.astype(np.int32))
    return np.ascontiguousarray(a.astype(np.int64))


def to_dev(a: np.ndarray) -> torch.Tensor:
    dev = require_cuda()
    return torch.from_numpy(a).to(dev, non_blocking=False)


def to_host(t: torch.Tensor) -> np.ndarray:
    return t.cpu().numpy()
