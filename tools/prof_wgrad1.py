"""Time the 1x1 weight gradient (sg3_modconv_wgrad, split-K tcgen05 GEMM) at the StyleGAN3-R 1024^2 layer shapes against its HBM floor
(dy and x read once) and the tensor-core time:  python tools/prof_wgrad1.py [N]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sg3_b200  # noqa: F401
from sg3_b200 import capi
from oracle import sg3_oracle as orc

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
_, specs = orc.layer_specs(1024, channel_base=65536, channel_max=1024, conv_kernel=1, use_radial_filters=True)
tot = floor = 0.0
print('| layer | I | O | H | ms | TFLOP/s | TB/s | of HBM floor |\n|---|---|---|---|---|---|---|---|')
for sp in specs:
    I, O, H = sp['in_channels'], sp['out_channels'], sp['in_size']
    if (H * H) % 4:
        continue
    ldw = (I + 31) // 32 * 32
    x = torch.randn(N, I, H, H, device='cuda')
    dy = torch.randn(N, O, H, H, device='cuda')
    dw = torch.zeros(N, O, ldw, device='cuda')

    def run():
        rc = capi.lib().sg3_modconv_wgrad(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), N, I, O, H, H, ldw, capi.stream_ptr(x.device))
        assert rc == 0
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    flops = 2.0 * N * O * I * H * H
    byts = 4.0 * N * (I + O) * H * H
    fl = max(byts / 6.54e12, flops / 1.1e15) * 1e3
    tot += ms
    floor += fl
    print(f"| {sp['name']} | {I} | {O} | {H} | {ms:.3f} | {flops / ms / 1e9:.0f} | {byts / ms / 1e9:.2f} | {fl / ms:.2f} |", flush=True)
print(f'\ntotal N={N}: {tot:.2f} ms, floor {floor:.2f} ms')
