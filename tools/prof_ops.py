"""Achieved HBM GB/s of the streaming ops (bias_act, upfirdn2d, filtered_lrelu_act): python tools/prof_ops.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import sg3_b200
from sg3_b200 import bias_act, upfirdn2d


def timeit(fn, iters=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


dev = 'cuda'
rows = []
for dt in (torch.float32, torch.float16):
    x = torch.randn(8, 256, 512, 512, device=dev, dtype=dt); b = torch.randn(256, device=dev, dtype=dt)
    ms = timeit(lambda: bias_act.bias_act(x, b, act='lrelu'))
    rows.append((f'bias_act lrelu {list(x.shape)} {str(dt)[6:]}', 2 * x.numel() * x.element_size(), ms))
    ms = timeit(lambda: bias_act.bias_act(x, b, act='linear', clamp=256))
    rows.append((f'bias_act linear+clamp {list(x.shape)} {str(dt)[6:]}', 2 * x.numel() * x.element_size(), ms))
f = upfirdn2d.setup_filter([1, 3, 3, 1], device=dev)
f12 = upfirdn2d.setup_filter(np.hanning(14)[1:-1], device=dev)
f2d = upfirdn2d.setup_filter(np.outer(np.hanning(14)[1:-1], np.hanning(14)[1:-1]), device=dev, separable=False)
for name, fn, shp in [
    ('filter2d 4-tap sep', lambda x: upfirdn2d.filter2d(x, f), (8, 64, 1024, 1024)),
    ('upsample2d x2 4-tap sep', lambda x: upfirdn2d.upsample2d(x, f, up=2), (8, 64, 512, 512)),
    ('downsample2d /2 4-tap sep', lambda x: upfirdn2d.downsample2d(x, f, down=2), (8, 64, 1024, 1024)),
    ('upfirdn2d up2 12-tap sep (flrelu ref path)', lambda x: upfirdn2d.upfirdn2d(x, f12, up=2, padding=[11, 10, 11, 10], gain=4), (4, 64, 1044, 1044)),
    ('upfirdn2d down2 12-tap sep (flrelu ref path)', lambda x: upfirdn2d.upfirdn2d(x, f12, down=2), (4, 64, 2098, 2098)),
    ('upfirdn2d down2 12x12 dense (flrelu ref path)', lambda x: upfirdn2d.upfirdn2d(x, f2d, down=2), (4, 64, 2098, 2098)),
]:
    x = torch.randn(*shp, device=dev)
    y = fn(x)
    ms = timeit(lambda: fn(x))
    rows.append((f'{name} {list(shp)} -> {list(y.shape)}', 4 * (x.numel() + y.numel()), ms))
print('| op | algorithmic MB | ms | GB/s | of 6540 |\n|---|---:|---:|---:|---:|')
for n, byts, ms in rows:
    print(f'| {n} | {byts / 1e6:.0f} | {ms:.3f} | {byts / ms / 1e6:.0f} | {byts / ms / 1e6 / 6540:.2f} |')
