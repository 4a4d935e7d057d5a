"""Can the reference itself (baseline/_ref, the unmodified JVBSouza/descriptools install) run on this box?

Times its Numba CPU-jit twins (slope.py:8, flowhand.py:127, gfi.py:45,118) on a conditioned synthetic DEM and tries
its Numba-CUDA entry points (sloper slope.py:96, flow_hand_index flowhand.py:242, gfi_calculator gfi.py:150) on the
GPU.  D8 and flow accumulation do not exist in the reference: those two rasters come from this
repo's own device path, exactly as the reference's example loads them from files (example.py:36,39).

    python scripts/ref_probe.py [N=2048]   ->  JSON lines
"""
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "baseline", "_ref"))

import numpy as np


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    px, thr = 12.5, 2000
    out = {"n": n}
    try:
        from numba import cuda
        out["numba_cuda_available"] = bool(cuda.is_available())
        if out["numba_cuda_available"]:
            d = cuda.get_current_device()
            out["numba_device"] = {"name": d.name.decode() if isinstance(d.name, bytes) else str(d.name), "cc": list(d.compute_capability)}
    except Exception as ex:
        out["numba_cuda_error"] = repr(ex)[:300]
    print(json.dumps(out), flush=True)

    import torch
    from descriptools_b200 import device
    dem_t = device.conditioned_dem(n, n)
    slope_t, d8_t = device.slope_d8(dem_t, px)
    acc_t = device.flow_accumulation(d8_t, dtype=torch.int32)
    dem = dem_t.cpu().numpy()
    d8 = d8_t.cpu().numpy()
    acc = acc_t.cpu().numpy().astype(np.int64)
    river = (acc > thr).astype(np.int8)
    ours = device.hand(d8_t, dem_t, px, acc=acc_t, river_threshold=thr, gfi_params=(0.4, 0.1, px), idx_dtype=torch.int32)
    torch.cuda.synchronize()

    import descriptools.slope as rslope
    import descriptools.flowhand as rflow
    import descriptools.gfi as rgfi

    # ---- CPU-jit twins (1 thread: no prange in the reference) ----
    def cpu_chain():
        t = {}
        t0 = time.perf_counter(); s = rslope.slope_sequential_jit(dem, px); t["slope"] = time.perf_counter() - t0
        t0 = time.perf_counter(); fd, idx = rflow.fdist_indexes_sequential_jit(d8, river, px); t["fdist_idx"] = time.perf_counter() - t0
        t0 = time.perf_counter(); hand = rflow.hand_calculator(dem, idx); t["hand"] = time.perf_counter() - t0
        t0 = time.perf_counter(); ra = rgfi.river_accumulation(np.ascontiguousarray(acc), np.ascontiguousarray(idx.astype(np.int64))); t["river_acc"] = time.perf_counter() - t0
        t0 = time.perf_counter(); g = rgfi.geomorphic_flood_index_sequential_jit(hand, ra, 0.4, 0.1, px); t["gfi"] = time.perf_counter() - t0
        return t, (s, fd, idx, hand, g)
    try:
        cpu_chain()  # JIT warm-up
        t, res = cpu_chain()
        tot = sum(t.values())
        s, fd, idx, hand, g = res
        chk = {"slope_equal": bool(np.array_equal(s.astype(np.float32), slope_t.cpu().numpy())),
               "idx_equal": bool(np.array_equal(np.asarray(idx, dtype=np.int64), ours["idx"].cpu().numpy().astype(np.int64))),
               "hand_equal": bool(np.array_equal(np.asarray(hand, dtype=np.float32), ours["hand"].cpu().numpy()))}
        print(json.dumps({"reference_cpu_jit": {"seconds": t, "total_s": tot, "mcells_s": n * n / tot / 1e6, "cores": 1,
                                                "note": "slope+HAND+GFI only (the reference has no D8 / flow accumulation)", "vs_ours": chk}}), flush=True)
    except Exception as ex:
        print(json.dumps({"reference_cpu_jit_error": repr(ex)[:400]}), flush=True)

    # ---- the reference's own Numba-CUDA path on this GPU ----
    try:
        def gpu_chain():
            t = {}
            t0 = time.perf_counter(); s = rslope.sloper(dem, px); t["sloper"] = time.perf_counter() - t0
            t0 = time.perf_counter(); fd, idx, hand = rflow.flow_hand_index(dem, d8, river, px); t["flow_hand_index"] = time.perf_counter() - t0
            t0 = time.perf_counter(); g = rgfi.gfi_calculator(hand, acc, idx, 0.4, 0.1, px); t["gfi_calculator"] = time.perf_counter() - t0
            return t, (s, fd, idx, hand, g)
        gpu_chain()
        t, res = gpu_chain()
        tot = sum(t.values())
        s, fd, idx, hand, g = res
        chk = {"slope_equal": bool(np.array_equal(s.astype(np.float32), slope_t.cpu().numpy())),
               "idx_equal": bool(np.array_equal(np.asarray(idx, dtype=np.int64), ours["idx"].cpu().numpy().astype(np.int64))),
               "hand_equal": bool(np.array_equal(np.asarray(hand, dtype=np.float32), ours["hand"].cpu().numpy())),
               "fdist_close": bool(np.allclose(fd, ours["fdist"].cpu().numpy(), rtol=1e-5)),
               "gfi_close": bool(np.allclose(g, ours["gfi"].cpu().numpy().astype(np.float64), rtol=1e-5, atol=1e-6))}
        print(json.dumps({"reference_numba_cuda": {"seconds": t, "total_s": tot, "mcells_s": n * n / tot / 1e6,
                                                   "note": "host arrays in/out per call, as the reference's API does", "vs_ours": chk}}), flush=True)
    except Exception as ex:
        print(json.dumps({"reference_numba_cuda_error": repr(ex)[:600]}), flush=True)


if __name__ == "__main__":
    main()
