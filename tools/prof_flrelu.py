"""One filtered_lrelu launch at a BASELINE-size layer, for ncu: python tools/prof_flrelu.py [L11|L10|TL12|all|Tall] [N] [bwd]  (T prefix = config T)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sg3_b200
from oracle import sg3_oracle as orc
which = sys.argv[1] if len(sys.argv) > 1 else 'L11'
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfgT = which.startswith('T')
if cfgT:
    which = which[1:]
    _, specs = orc.layer_specs(1024, channel_base=32768, channel_max=512, conv_kernel=3, use_radial_filters=False)
else:
    _, specs = orc.layer_specs(1024, channel_base=65536, channel_max=1024, conv_kernel=1, use_radial_filters=True)
if which == 'all':
    import subprocess
    for i in range(len(specs) - 1):
        subprocess.run([sys.executable, __file__, ('T' if cfgT else '') + f'L{i}', str(N)])
    sys.exit(0)
sp = specs[int(which[1:])]
C, size = sp['out_channels'], sp['in_size'] + sp['conv_kernel'] - 1
x = torch.randn(N, C, size, size, device='cuda') * 2
b = torch.randn(C, device='cuda')
fu = torch.from_numpy(sp['up_filter']).cuda(); fd = torch.from_numpy(sp['down_filter']).cuda()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for it in range(4):
    if it == 3: ev[0].record()
    y = sg3_b200.filtered_lrelu.filtered_lrelu(x, fu, fd, b, up=sp['up'], down=sp['down'], padding=sp['padding'], clamp=256)
ev[1].record(); torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1])
byts = 4 * (x.numel() + y.numel())
print(f'{sp["name"]} up{sp["up"]} N={N} C={C} {size}->{y.shape[-1]}: {ms:.3f} ms, {byts / ms / 1e6:.1f} GB/s, {y.numel() / ms / 1e6:.1f} Gpix/s out')
if len(sys.argv) > 3 and sys.argv[3] == 'bwd':
    xg = x.clone().requires_grad_(True)
    for it in range(3):
        y = sg3_b200.filtered_lrelu.filtered_lrelu(xg, fu, fd, b, up=sp['up'], down=sp['down'], padding=sp['padding'], clamp=256)
        dy = torch.randn_like(y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        (dx,) = torch.autograd.grad(y, xg, dy)
        e1.record(); torch.cuda.synchronize()
    print(f'   backward (sign read): {e0.elapsed_time(e1):.3f} ms   forward with sign write: see below')
    e0.record()
    y = sg3_b200.filtered_lrelu.filtered_lrelu(xg, fu, fd, b, up=sp['up'], down=sp['down'], padding=sp['padding'], clamp=256)
    e1.record(); torch.cuda.synchronize()
    print(f'   forward (sign write): {e0.elapsed_time(e1):.3f} ms')
