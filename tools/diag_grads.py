import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
import sg3_b200
from sg3_b200 import networks, modulated_conv
from conftest import golden, rel_err
sg3_b200.filtered_lrelu._quiet_fallback = True
name = sys.argv[1] if len(sys.argv) > 1 else 'tinyR'
cfg = dict(z_dim=64, c_dim=0, w_dim=64, img_resolution=64, img_channels=3, channel_base=2048, channel_max=32,
           conv_kernel=1 if name == 'tinyR' else 3, use_radial_filters=(name == 'tinyR'))
torch.manual_seed(0)
G = networks.Generator(**cfg).cuda()
g = golden('tiny.npz')
modulated_conv.set_math('fp32')
ws = torch.from_numpy(g.z[f'{name}/ws']).cuda()
for impl in ['cuda', 'ref']:
    orig = sg3_b200.filtered_lrelu.filtered_lrelu
    if impl == 'ref':
        sg3_b200.filtered_lrelu.filtered_lrelu = lambda *a, **k: orig(*a, **{**k, 'impl': 'ref'})
    img = G.synthesis(ws, noise_mode='const', force_fp32=True)
    loss = (img * torch.from_numpy(g.z[f'{name}/tgt']).cuda()).mean()
    params = dict(G.synthesis.named_parameters())
    grads = torch.autograd.grad(loss, list(params.values()), allow_unused=True)
    sg3_b200.filtered_lrelu.filtered_lrelu = orig
    print('impl', impl, 'loss', float(loss), float(g.z[f'{name}/loss']))
    for (k, _), gr in zip(params.items(), grads):
        key = f'{name}/grad/{k}'
        if key in g.z.files and np.abs(g.z[key]).max() > 0:
            e = rel_err(gr.cpu().numpy(), g.z[key])
            if e > 3e-4: print(f'  {k:40s} {e:.2e}')
