"""1024^2 synthesis forward with fp32 activations (force_fp32=True) vs fp16 layers (reference default): python tools/prof_fp16.py [batch] [R|T]
(config T also times the fp16 layers with their 3x3 convs upcast to the TF32 kernel, the path before the fp16 3x3 kernel existed)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sg3_b200
from sg3_b200 import networks, capi

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
CFG = sys.argv[2] if len(sys.argv) > 2 else 'R'
torch.manual_seed(0)
if CFG == 'T':
    G = networks.Generator(z_dim=512, c_dim=0, w_dim=512, img_resolution=1024, img_channels=3, channel_base=32768, channel_max=512,
                           conv_kernel=3, use_radial_filters=False).eval().requires_grad_(False).cuda()
else:
    G = networks.Generator(z_dim=512, c_dim=0, w_dim=512, img_resolution=1024, img_channels=3, channel_base=65536, channel_max=1024,
                           conv_kernel=1, use_radial_filters=True).eval().requires_grad_(False).cuda()
from sg3_b200 import modulated_conv as _mc
_orig = _mc.modulated_conv2d


def _upcast3(x, w, s, **kw):
    if x.dtype == torch.float16 and w.shape[-1] == 3:
        return _orig(x.float(), w, s, **kw).to(x.dtype)
    return _orig(x, w, s, **kw)
ws = G.mapping(torch.randn(B, 512, device='cuda'), None)
outs = {}
arms = [('force_fp32=True', dict(force_fp32=True)), ('fp16 layers (default)', dict())]
if CFG == 'T':
    arms.append(('fp16 layers, 3x3 upcast', dict()))
for name, kw in arms:
    networks.modulated_conv2d = _upcast3 if 'upcast' in name else _orig
    with torch.no_grad():
        for _ in range(3): img = G.synthesis(ws, noise_mode='const', **kw)
        n0 = capi.lib().sg3_launch_count()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5): img = G.synthesis(ws, noise_mode='const', **kw)
        torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 5 * 1e3
    outs[name] = img
    print(f'{name:24s}: {ms:8.2f} ms / {B} images = {B / ms * 1e3:7.1f} img/s, {(capi.lib().sg3_launch_count() - n0) // 5} sg3 launches')
a, b = outs['force_fp32=True'], outs['fp16 layers (default)']
print('max |fp16 - fp32| / max |fp32| =', ((a - b).abs().max() / a.abs().max()).item())
