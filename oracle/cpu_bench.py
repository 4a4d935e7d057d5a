"""CPU timing of the oracle port over the headline chain -- run as a separate process by bench.py.

Test infrastructure (see oracle/__init__.py).  A process of its own, without torch: a second OpenMP runtime in
the process (torch bundles one) and torchrun's OMP_NUM_THREADS=1 both distort the timing of the OpenMP loops.

    python oracle/cpu_bench.py ROWS COLS STEPS WARMUP PX THR N B   ->  one JSON line
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rows, cols, steps, warmup = (int(v) for v in sys.argv[1:5])
    px, thr, n, b = float(sys.argv[5]), int(sys.argv[6]), float(sys.argv[7]), float(sys.argv[8])
    import numpy as np

    import oracle

    oracle.build()
    cores = oracle.set_threads(os.cpu_count() or 1)
    dem = oracle.conditioned_dem(rows, cols)  # same recipe as the GPU workload (dtb-synth-v1), conditioned on the host

    def chain():
        t0 = time.perf_counter()
        slope, d8 = oracle.slope_d8(dem, px)
        acc, _ = oracle.flow_accumulation(d8)
        river = (acc > thr).astype(np.int8)
        fdist, idx, hand = oracle.flow_hand_index(dem, d8, river, px)
        oracle.gfi(hand, acc, idx, n, b, px)
        return time.perf_counter() - t0

    for _ in range(warmup):
        chain()
    t = [chain() for _ in range(steps)]
    print(json.dumps({"seconds": t, "cores": cores, "cells": rows * cols}))


if __name__ == "__main__":
    main()
