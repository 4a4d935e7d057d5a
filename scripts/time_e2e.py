import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from descriptools_b200 import device, pipeline
rows = cols = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dem = device.conditioned_dem(rows, cols)
dts = pipeline.output_dtypes(np.float32, rows * cols)
dem_host = torch.empty((rows, cols), dtype=torch.float32, pin_memory=True); dem_host.copy_(dem)
pinned = {k: torch.empty((rows, cols), dtype=getattr(torch, np.dtype(v).name), pin_memory=True) for k, v in dts.items()}
del dem; device.workspace.release(); torch.cuda.empty_cache()
for i in range(5):
    torch.cuda.synchronize(); t = time.perf_counter()
    pipeline.pipeline(dem_host, 12.5, 128000, 0.4, 0.1, pinned_out=pinned, chunks=chunks)
    torch.cuda.synchronize(); print(f"e2e {i}: {(time.perf_counter()-t)*1e3:.1f} ms", flush=True)
