"""One upfirdn2d case a few times (for ncu):  python tools/prof_upfirdn.py up2|down2|filter4|up4tap|down4tap [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import sg3_b200  # noqa: F401
from sg3_b200 import upfirdn2d

case = sys.argv[1] if len(sys.argv) > 1 else 'down2'
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = 'cuda'
f12 = upfirdn2d.setup_filter(np.hanning(14)[1:-1], device=dev)
f4 = upfirdn2d.setup_filter([1, 3, 3, 1], device=dev)
if case == 'up2':
    x = torch.randn(4, 64, 1044, 1044, device=dev)
    fn = lambda: upfirdn2d.upfirdn2d(x, f12, up=2, padding=[11, 10, 11, 10], gain=4)
elif case == 'down2':
    x = torch.randn(4, 64, 2098, 2098, device=dev)
    fn = lambda: upfirdn2d.upfirdn2d(x, f12, down=2)
elif case == 'filter4':
    x = torch.randn(8, 64, 1024, 1024, device=dev)
    fn = lambda: upfirdn2d.filter2d(x, f4)
elif case == 'up4tap':
    x = torch.randn(8, 64, 512, 512, device=dev)
    fn = lambda: upfirdn2d.upsample2d(x, f4, up=2)
else:
    x = torch.randn(8, 64, 1024, 1024, device=dev)
    fn = lambda: upfirdn2d.downsample2d(x, f4, down=2)
for _ in range(iters):
    y = fn()
torch.cuda.synchronize()
print(case, tuple(y.shape))
