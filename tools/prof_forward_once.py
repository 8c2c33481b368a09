"""One synthesis forward at a small batch, nothing else -- the ncu target for per-launch DRAM traffic of the filtered_lrelu kernels:
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file out.csv python tools/prof_forward_once.py [R|T] [batch]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sg3_b200
from sg3_b200 import networks

cfg = sys.argv[1] if len(sys.argv) > 1 else 'R'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
kw = dict(R=dict(channel_base=65536, channel_max=1024, conv_kernel=1, use_radial_filters=True),
          T=dict(channel_base=32768, channel_max=512, conv_kernel=3, use_radial_filters=False))[cfg]
torch.manual_seed(0)
G = networks.Generator(z_dim=512, c_dim=0, w_dim=512, img_resolution=1024, img_channels=3, **kw).eval().requires_grad_(False).cuda()
sg3_b200.filtered_lrelu._quiet_fallback = True
with torch.no_grad():
    ws = G.mapping(torch.randn(B, 512, device='cuda'), None)
    img = G.synthesis(ws, noise_mode='const', force_fp32=True)
torch.cuda.synchronize()
print(cfg, B, tuple(img.shape))
