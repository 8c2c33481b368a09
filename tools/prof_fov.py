"""FOV expansion of one 1024^2 frame (9 views): nine sequential batch-1 synthesis calls (the reference's procedure,
utils/fov_expansion.py:14-31) vs sg3_b200.fov.Expander (one batch-9 call).  python tools/prof_fov.py [R|T]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sg3_b200
from sg3_b200 import fov, networks
cfg = networks.CONFIG_T if (len(sys.argv) > 1 and sys.argv[1] == 'T') else networks.CONFIG_R
torch.manual_seed(0)
G = networks.Generator(z_dim=512, c_dim=0, w_dim=512, img_resolution=1024, img_channels=3, **cfg).eval().requires_grad_(False).cuda()
ws = G.mapping(torch.randn(1, 512, device='cuda'), None)
lt = np.eye(3)
views = fov.view_transforms(1024, 100, 100, 100, 100)
ex = fov.Expander(G)


def sequential():
    out = []
    for t in views:
        G.synthesis.input.transform = torch.from_numpy(lt @ t).float().cuda()
        with torch.no_grad():
            out.append(G.synthesis(ws))
    return out


def timed(fn):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5


a = timed(sequential)
b = timed(lambda: ex.generate_expanded_image(ws=ws, landmark_t=lt, pixels_right=100, pixels_left=100, pixels_top=100, pixels_bottom=100))
print(f'config {"T" if cfg is networks.CONFIG_T else "R"} 1024^2, fp16 layers (reference default), 9 views of one frame: '
      f'sequential batch-1 calls {a:.2f} ms, one batch-9 call + paste {b:.2f} ms  ({a / b:.2f}x)')
