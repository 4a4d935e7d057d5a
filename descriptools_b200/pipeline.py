"""pipeline.py -- the headline chain slope -> D8 -> flow accumulation -> HAND -> GFI.

The reference runs every descriptor as an independent host->device->host round trip
(Example/example.py:59-91) and takes D8 / accumulation rasters from disk (example.py:36,39).
Here the chain stays on the device: three C-ABI calls (dtb_slope_d8, dtb_flowacc, dtb_hand with
the fused HAND/GFI epilogue) over buffers that never leave HBM.

    run_device(dem_cuda_tensor, ...) -> dict of CUDA tensors        (what bench.py times as `value`)
    pipeline(dem_numpy, ...)         -> dict of NumPy arrays        (host in / host out, `e2e`)
"""
from __future__ import annotations

import numpy as np
import torch

from . import device
from ._convert import dem_to_native

STAGE_OUTPUTS = ("slope", "d8", "acc", "fdist", "idx", "hand", "gfi")


def run_device(dem: torch.Tensor, px: float, river_threshold: int, n_gfi: float = 0.4, scale_factor: float = 0.1,
               size: float | None = None, max_moves: int = 0) -> dict:
    """slope(%) f32, d8 u8, acc i32/i64, fdist f32, idx i32/i64, hand (DEM dtype), gfi f32.

    river cells are `acc > river_threshold` (example.py:52); GFI uses `size` = px by default
    (example.py:87).  All tensors stay on the device.
    """
    size = px if size is None else size
    rows, cols = dem.shape
    int_dt = torch.int32 if rows * cols < 2**31 else torch.int64
    slope, d8 = device.slope_d8(dem, px)
    acc = device.flow_accumulation(d8, dtype=int_dt, nodata_fill=-100)
    out = device.hand(d8, dem, px, acc=acc, river_threshold=river_threshold, max_moves=max_moves,
                      gfi_params=(n_gfi, scale_factor, size), idx_dtype=int_dt)
    out.update(slope=slope, d8=d8, acc=acc)
    return out


def pipeline(dem, px: float, river_threshold: int, n_gfi: float = 0.4, scale_factor: float = 0.1,
             size: float | None = None, outputs=STAGE_OUTPUTS, pinned_out: dict | None = None) -> dict:
    """Host-side chain: NumPy DEM in, NumPy rasters out (dtypes as the reference returns them on the
    device side: slope/fdist/gfi float32, d8 uint8, acc/idx int32 or int64, hand in the DEM's dtype).

    `pinned_out` may hold pre-allocated pinned torch CPU tensors keyed by output name (the benchmark
    reuses them across steps); otherwise pageable arrays are returned.
    """
    dev = device.require_cuda()
    if isinstance(dem, torch.Tensor):
        dem_t = dem.to(dev, non_blocking=True)
    else:
        dem_t = torch.from_numpy(dem_to_native(dem)).to(dev, non_blocking=True)
    res = run_device(dem_t, px, river_threshold, n_gfi, scale_factor, size)
    host = {}
    for name in outputs:
        t = res[name]
        if pinned_out is not None and name in pinned_out:
            pinned_out[name].copy_(t, non_blocking=True)
            host[name] = pinned_out[name].numpy()
        else:
            host[name] = t.cpu().numpy()
    torch.cuda.current_stream().synchronize()
    return host


def output_dtypes(dem_dtype: np.dtype, n_cells: int) -> dict:
    it = np.int32 if n_cells < 2**31 else np.int64
    return dict(slope=np.float32, d8=np.uint8, acc=it, fdist=np.float32, idx=it, hand=np.dtype(dem_dtype), gfi=np.float32)
