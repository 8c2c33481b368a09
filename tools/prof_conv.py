"""Time the modulated-conv contraction (tcgen05 path) at the BASELINE layer shapes: python tools/prof_conv.py [N] [R|T]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sg3_b200
from sg3_b200 import capi
from oracle import sg3_oracle as orc
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
CFG = sys.argv[2] if len(sys.argv) > 2 else 'R'
ONLY = sys.argv[3] if len(sys.argv) > 3 else None
if CFG == 'R':
    _, specs = orc.layer_specs(1024, channel_base=65536, channel_max=1024, conv_kernel=1, use_radial_filters=True)
else:
    _, specs = orc.layer_specs(1024, channel_base=32768, channel_max=512, conv_kernel=3, use_radial_filters=False)
tot = 0; tot_ideal = 0
for sp in specs:
    if ONLY and not sp['name'].startswith(ONLY + '_'): continue
    I, O, H = sp['in_channels'], sp['out_channels'], sp['in_size']
    k = sp.get('conv_kernel', 1)
    if not sp.get('is_torgb', False) and CFG == 'T': k = 3
    if sp.get('is_torgb', False): k = 1
    pad = k - 1
    OH = H + pad
    ldw = (I + 31) // 32 * 32
    x = torch.randn(N, I, H, H, device='cuda'); w = torch.randn(N, k * k, O, ldw, device='cuda'); y = torch.empty(N, O, OH, OH, device='cuda')
    def run():
        rc = capi.lib().sg3_modconv_fwd(x.data_ptr(), w.data_ptr(), y.data_ptr(), N, I, O, H, H, k, pad, ldw, 1, 0, capi.stream_ptr(x.device))
        assert rc == 0
    for _ in range(3): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    flops = 2.0 * N * O * I * k * k * OH * OH; byts = 4.0 * N * (I * H * H + O * OH * OH) + 4.0 * N * O * ldw * k * k
    ideal = max(flops / 1.1e15, byts / 6.54e12) * 1e3
    tot += ms; tot_ideal += ideal
    print(f"{sp['name']:16s} I={I:4d} O={O:4d} H={H:4d}: {ms*1e3:8.1f} us  {flops/ms/1e9:7.1f} TFLOP/s  {byts/ms/1e9:6.2f} TB/s  ideal {ideal*1e3:7.1f} us ({100*ideal/ms:4.0f}%)")
print(f'total {tot:.3f} ms for N={N} ({tot/N:.3f} ms/img), ideal {tot_ideal:.3f} ms')
