"""Small-batch latency of the R-1024 synthesis forward, eager vs CUDA-graph replay: python tools/prof_latency.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sg3_b200
from sg3_b200 import networks

torch.manual_seed(0)
G = networks.Generator(z_dim=512, c_dim=0, w_dim=512, img_resolution=1024, img_channels=3, channel_base=65536, channel_max=1024,
                       conv_kernel=1, use_radial_filters=True).eval().requires_grad_(False).cuda()
print('| batch | eager ms | graph ms | eager img/s | graph img/s | max abs diff |\n|---:|---:|---:|---:|---:|---:|')
for B in (1, 2, 4, 8):
    ws = G.mapping(torch.randn(B, 512, device='cuda'), None)
    with torch.no_grad():
        for _ in range(3): ref = G.synthesis(ws, noise_mode='const', force_fp32=True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10): ref = G.synthesis(ws, noise_mode='const', force_fp32=True)
        torch.cuda.synchronize(); eager = (time.perf_counter() - t0) / 10 * 1e3
    gs = networks.GraphedSynthesis(G.synthesis, ws)
    for _ in range(3): out = gs(ws)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): out = gs(ws)
    torch.cuda.synchronize(); graph = (time.perf_counter() - t0) / 10 * 1e3
    print(f'| {B} | {eager:.2f} | {graph:.2f} | {B / eager * 1e3:.1f} | {B / graph * 1e3:.1f} | {(out - ref).abs().max().item():.1e} |')
