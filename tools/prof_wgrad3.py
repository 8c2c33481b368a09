"""Time the 3x3 weight gradient at the StyleGAN3-T 1024^2 layer shapes: the tcgen05 kernel (sg3_modconv_wgrad3) next to the
library call it replaces (torch.nn.grad.conv2d_weight, groups = N, TF32 allowed = what conv2d_gradfix.py:153-174 runs).

    python tools/prof_wgrad3.py [N] [first layer prefix]
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sg3_b200  # noqa: F401
from sg3_b200 import modulated_conv as mc
from oracle import sg3_oracle as orc

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
ONLY = sys.argv[2] if len(sys.argv) > 2 else None
_, specs = orc.layer_specs(1024, channel_base=32768, channel_max=512, conv_kernel=3, use_radial_filters=False)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


tot = [0.0, 0.0]
print(f'| layer | I | O | H | ours ms | TFLOP/s | library ms | ratio | max rel diff |\n|---|---|---|---|---|---|---|---|---|')
for sp in specs:
    if sp.get('is_torgb', False):
        continue
    if ONLY and not sp['name'].startswith(ONLY):
        continue
    I, O, H = sp['in_channels'], sp['out_channels'], sp['in_size']
    OH = H + 2
    x = torch.randn(N, I, H, H, device='cuda')
    dy = mc.empty_row_pitched([N, O, OH, OH], torch.float32, 'cuda')
    dy.copy_(torch.randn(N, O, OH, OH, device='cuda'))
    dyc = dy.contiguous()
    ours = timed(lambda: mc.conv3x3_weight_grad(x, dy, 2))

    def lib():
        with torch.backends.cudnn.flags(allow_tf32=True):
            return torch.nn.grad.conv2d_weight(x.reshape(1, N * I, H, H), [N * O, I, 3, 3], dyc.reshape(1, N * O, OH, OH), padding=2, groups=N)
    ref = timed(lib)
    a, b = mc.conv3x3_weight_grad(x, dy, 2), lib().reshape(N, O, I, 3, 3)
    diff = float((a - b).abs().max() / b.abs().max())
    flops = 2.0 * N * O * I * 9 * OH * OH
    tot[0] += ours
    tot[1] += ref
    print(f"| {sp['name']} | {I} | {O} | {H} | {ours:.3f} | {flops / ours / 1e9:.0f} | {ref:.3f} | {ref / ours:.2f} | {diff:.1e} |", flush=True)
print(f'\ntotal N={N}: ours {tot[0]:.2f} ms, library {tot[1]:.2f} ms ({tot[1] / tot[0]:.2f}x)')
