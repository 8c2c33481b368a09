"""Debug helper for the tcgen05 contraction: structured inputs, NaN-prefilled output."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sg3_b200
from sg3_b200 import capi

def run(N, I, O, H, W, x, wm):
    ldw = (I + 31) // 32 * 32
    wpad = torch.zeros(N, O, ldw, device='cuda'); wpad[:, :, :I] = wm
    y = torch.full((N, O, H, W), float('nan'), device='cuda')
    rc = capi.lib().sg3_modconv_fwd(x.data_ptr(), wpad.data_ptr(), y.data_ptr(), N, I, O, H, W, 1, 0, ldw, 1, 0, capi.stream_ptr(x.device))
    torch.cuda.synchronize()
    return rc, y

N, I, O, H, W = 1, 32, 16, 16, 16
x = torch.ones(N, I, H, W, device='cuda'); wm = torch.ones(N, O, I, device='cuda')
rc, y = run(N, I, O, H, W, x, wm)
print('rc', rc, 'nan frac', float(torch.isnan(y).float().mean()), 'uniq', torch.unique(y[~torch.isnan(y)])[:10].tolist())
# x = pixel index, w = delta on i=0  -> y[o][p] = p
x = torch.zeros(N, I, H, W, device='cuda'); x[0, 0] = torch.arange(H * W, device='cuda').float().reshape(H, W)
wm = torch.zeros(N, O, I, device='cuda'); wm[0, :, 0] = 1
rc, y = run(N, I, O, H, W, x, wm)
print('pix test: y[0,0].flatten()[:40]', y[0, 0].flatten()[:40].tolist())
print('pix test: y[0,3].flatten()[100:110]', y[0, 3].flatten()[100:110].tolist())
# x = 1 on channel i only, w[o][i] = 100*o + i -> y[o][p] = 100*o + i
for i in (0, 1, 7, 8, 31):
    x = torch.zeros(N, I, H, W, device='cuda'); x[0, i] = 1
    wm = (100 * torch.arange(O, device='cuda').float()[:, None] + torch.arange(I, device='cuda').float()[None, :])[None]
    rc, y = run(N, I, O, H, W, x, wm.contiguous())
    print('chan', i, 'y[0,:,0,0]', y[0, :, 0, 0].tolist())
