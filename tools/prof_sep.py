import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import sg3_b200
from sg3_b200 import upfirdn2d
f12 = upfirdn2d.setup_filter(np.hanning(14)[1:-1], device='cuda')
mode = sys.argv[1]
if mode == 'up':
    x = torch.randn(4, 64, 1044, 1044, device='cuda'); fn = lambda: upfirdn2d.upfirdn2d(x, f12, up=2, padding=[11, 10, 11, 10], gain=4)
else:
    x = torch.randn(4, 64, 2098, 2098, device='cuda'); fn = lambda: upfirdn2d.upfirdn2d(x, f12, down=2)
for _ in range(4): y = fn()
torch.cuda.synchronize()
