import sys, torch
sys.path.insert(0, ".")
import bench
from descriptools_b200 import device, pipeline
n = 10000
dem = device.conditioned_dem(n, n)
res = pipeline.run_device(dem, 12.5, 2000)
ms = bench._time_launches(lambda: device.downslope(dem, res["d8"], 12.5, 5.0), reps=5, warm=2)
print("downslope", n, round(ms, 3), "ms")
