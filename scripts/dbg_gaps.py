import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from descriptools_b200 import device, _lib
rows = cols = 40000
dem = device.conditioned_dem(rows, cols)
ev = lambda: torch.cuda.Event(enable_timing=True)
def step():
    e = [ev() for _ in range(6)]
    e[0].record()
    slope = torch.empty((rows, cols), dtype=torch.float32, device="cuda"); d8 = torch.empty((rows, cols), dtype=torch.uint8, device="cuda")
    e[1].record()
    device.check(device.lib.dtb_slope_d8(dem.data_ptr(), 0, rows, cols, 0, rows, 12.5, slope.data_ptr(), d8.data_ptr(), torch.cuda.current_stream().cuda_stream), "x")
    e[2].record()
    acc = device.flow_accumulation(d8, fuse_hand_threshold=128000)
    e[3].record()
    t0 = time.perf_counter()
    out = device.hand(d8, dem, 12.5, acc=acc, river_threshold=128000, gfi_params=(0.4, 0.1, 12.5), idx_dtype=torch.int32, entry_done=True)
    t1 = time.perf_counter()
    e[4].record()
    return e, (slope, d8, acc, out), t1 - t0
for prof in (False, True):
    _lib.profile_enable(prof)
    keep = None
    for _ in range(3):
        e, keep, th = step()
    torch.cuda.synchronize()
    t = time.perf_counter()
    es = []
    for _ in range(4):
        e, keep, th = step(); es.append(e)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t) / 4
    for e in es[-2:]:
        print("prof", prof, "alloc %.3f slope %.3f flowacc %.3f hand %.3f | wall/step %.3f ms host-hand-call %.3f ms" % (e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3]), e[3].elapsed_time(e[4]), wall * 1e3, th * 1e3))
    if prof:
        print({k: round(v[0] / 7, 3) for k, v in _lib.profile_collect().items()})
