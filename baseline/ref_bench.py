"""Timing of the UNMODIFIED reference (JVBSouza/descriptools, installed under baseline/_ref by
`pip install --target baseline/_ref /root/reference`) over the headline chain -- run as a separate process by bench.py.

    python baseline/ref_bench.py cpu ROWS COLS STEPS WARMUP PX THR N B   ->  one JSON line
    python baseline/ref_bench.py gpu ROWS COLS STEPS WARMUP PX THR N B   ->  one JSON line

cpu: the reference's compiled CPU-jit twins -- slope_sequential_jit (slope.py:8), fdist_indexes_sequential_jit
     (flowhand.py:127), hand_calculator (flowhand.py:414), geomorphic_flood_index_sequential_jit (gfi.py:45, which gathers
     the river accumulation itself).  They are serial Numba loops (no prange): ONE core, whatever the box has.
gpu: the reference's own public entry points, i.e. its Numba-CUDA kernels on this GPU with host arrays in and out per
     call, exactly as its API works: sloper (slope.py:96), flow_hand_index (flowhand.py:242), gfi_calculator (gfi.py:150).

D8 flow direction and flow accumulation do not exist in the reference (it loads both rasters from files,
Example/example.py:36,39).  For those two stages of the chain the only CPU implementation there is is this repo's
restatement (oracle/dt_oracle.c, OpenMP); its time is added and reported separately (`port_seconds`).
"""
import json
import os
import sys
import time
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(HERE, "_ref"))
warnings.simplefilter("ignore")


def main():
    mode = sys.argv[1]
    rows, cols, steps, warmup = (int(v) for v in sys.argv[2:6])
    px, thr, n, b = float(sys.argv[6]), int(sys.argv[7]), float(sys.argv[8]), float(sys.argv[9])
    import numpy as np

    import oracle

    oracle.build()
    cores = oracle.set_threads(os.cpu_count() or 1)
    dem = oracle.conditioned_dem(rows, cols)  # same recipe as the GPU workload (dtb-synth-v1), conditioned on the host

    import descriptools.flowhand as rflow
    import descriptools.gfi as rgfi
    import descriptools.slope as rslope
    import numba

    info = {"numba": numba.__version__, "reference": os.path.dirname(rslope.__file__)}
    if mode == "gpu":
        from numba import cuda

        if not cuda.is_available():
            print(json.dumps({"error": "numba.cuda is not available on this box"}))
            return
        d = cuda.get_current_device()
        info["device"] = d.name.decode() if isinstance(d.name, bytes) else str(d.name)

    def chain():
        t = {}
        t0 = time.perf_counter()
        if mode == "cpu":
            rslope.slope_sequential_jit(dem, px)
        else:
            rslope.sloper(dem, px)
        t["slope"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        _, d8 = oracle.slope_d8(dem, px)
        acc, _ = oracle.flow_accumulation(d8)
        t["port_d8_flowacc"] = time.perf_counter() - t0
        river = (acc > thr).astype(np.int8)
        acc64 = np.ascontiguousarray(acc.astype(np.int64))
        t0 = time.perf_counter()
        if mode == "cpu":
            fdist, idx = rflow.fdist_indexes_sequential_jit(d8, river, px)
            idx = np.ascontiguousarray(np.asarray(idx).astype(np.int64))
            hand = np.ascontiguousarray(rflow.hand_calculator(dem, idx))
        else:
            fdist, idx, hand = rflow.flow_hand_index(dem, d8, river, px)
            idx = np.ascontiguousarray(np.asarray(idx).astype(np.int64))
            hand = np.ascontiguousarray(hand)
        t["hand"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        if mode == "cpu":
            rgfi.geomorphic_flood_index_sequential_jit(hand, acc64, idx, n, b, px)
        else:
            rgfi.gfi_calculator(hand, acc64, idx, n, b, px)
        t["gfi"] = time.perf_counter() - t0
        return t

    for _ in range(max(warmup, 1)):  # at least one: the JIT compiles on first call
        chain()
    ts = [chain() for _ in range(steps)]
    tot = [sum(t.values()) for t in ts]
    ref_only = [sum(v for k, v in t.items() if not k.startswith("port_")) for t in ts]
    stage = {k: sum(t[k] for t in ts) / len(ts) for k in ts[0]}
    print(json.dumps({"seconds": tot, "reference_seconds": ref_only, "stage_seconds": stage, "cores": 1, "port_cores": cores,
                      "cells": rows * cols, **info}))


if __name__ == "__main__":
    main()
