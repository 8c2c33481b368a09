import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import sg3_b200
from sg3_b200 import networks, modulated_conv
import sg3_b200.networks as nw
from conftest import golden, rel_err
import test_network_gpu as T
g = golden('tiny.npz')
for name in ('tinyR', 'tinyT'):
    G, gg = T._build(sg3_b200, name)
    ws = torch.from_numpy(gg.z[name + '/ws']).cuda()
    ref = gg.z[name + '/img']
    modulated_conv.set_math('tf32')
    img = G.synthesis(ws, noise_mode='const', force_fp32=True)
    print(name, 'RN in stencil      ', rel_err(img.cpu().numpy(), ref))
    orig_mode = nw._math_mode
    nw._math_mode = lambda: 'no-round'
    orig_conv = nw.modulated_conv2d
    for c in (0.0, 2.0 ** -13, 2.0 ** -12, 3.5e-4, 2.0 ** -11):
        def conv(x, w, s, demodulate=True, padding=0, input_gain=None, _c=c):
            ig = input_gain * (1.0 + _c) if input_gain is not None else torch.tensor(1.0 + _c, device=x.device)
            return orig_conv(x=x, w=w, s=s, demodulate=demodulate, padding=padding, input_gain=ig)
        nw.modulated_conv2d = conv
        img = G.synthesis(ws, noise_mode='const', force_fp32=True)
        print(name, f'trunc + gain {c:.2e}', rel_err(img.cpu().numpy(), ref))
    nw.modulated_conv2d = orig_conv
    nw._math_mode = orig_mode
    modulated_conv.set_math(None)
