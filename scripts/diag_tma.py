import numpy as np, torch, sys
sys.path.insert(0,'.')
from descriptools_b200 import device
dem = torch.rand(256,256,device='cuda')*100+100
s,d = device.slope_d8(dem, 12.5)
torch.cuda.synchronize()
print('tma ok', float(s.max()))
