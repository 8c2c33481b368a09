"""Tensor-core FIR question: could the dense 12x12 down pass of filtered_lrelu run as a banded-Toeplitz
`tcgen05.mma` beside the SIMT stages?  Combines MEASURED MMA clocks (tools/tc_mma_bench FIRSWEEP lines) with the exact MAC
bookkeeping of the Toeplitz formulation, and measures the numerical error of TF32 / 3xTF32 operands on the real L11 filter.

    python tools/tc_fir_model.py gpurun_out/tc_mma_bench.log [simt_outputs_per_clk_per_sm]

Formulation (the only one whose operands can slide along both filter axes): A = activation rows from shared
memory, M = 128 output rows oy (row 2*oy + a of the plane with the matching parity: a row offset is a 16-byte step of the
K-major no-swizzle descriptor), K = a window of activation columns, B = the Toeplitz block T_a[x][ox] = F[a][x - 2*ox] of one
filter row a, N = outputs per MMA.  One output tile [128 oy][N ox] costs 12 filter rows x ceil((2*(N-1) + 12) / 8) MMAs of K = 8.
"""
import math
import re
import sys

import numpy as np

log = open(sys.argv[1]).read() if len(sys.argv) > 1 else ''
simt = float(sys.argv[2]) if len(sys.argv) > 2 else None
meas = {int(m.group(1)): float(m.group(2)) for m in re.finditer(r'FIRSWEEP N=\s*(\d+): no error issue [\d.]+ complete ([\d.]+)', log)}
print('| N (outputs per MMA) | K window | MMAs per [128 x N] tile | useful MAC share | measured clk / MMA (tensor floor N/2) | clk per tile | outputs / clk / SM (TF32) | 3xTF32 |')
print('|---|---|---|---|---|---|---|---|')
best = 0.0
for N in (16, 32, 48, 64, 128, 256):
    win = 2 * (N - 1) + 12
    ksteps = math.ceil(win / 8)
    mmas = 12 * ksteps
    useful = 12.0 / (ksteps * 8)
    clk = meas.get(N)
    if clk is None:
        print(f'| {N} | {win} -> {ksteps * 8} | {mmas} | {100 * useful:.1f} % | not measured | | | |')
        continue
    tile = mmas * clk
    opc = 128 * N / tile
    best = max(best, opc)
    print(f'| {N} | {win} -> {ksteps * 8} | {mmas} | {100 * useful:.1f} % | {clk:.1f} ({N // 2}) | {tile:.0f} | {opc:.2f} | {opc / 3:.2f} |')
print()
if meas:
    print(f'Best tensor-core rate for the down pass ALONE: {best:.2f} outputs/clk/SM with TF32 operands, {best / 3:.2f} with the 3xTF32 split that '
          f'fp32 parity needs' + (f'; the whole SIMT kernel (up + activation + down) runs at {simt:.2f} outputs/clk/SM.' if simt else '.'))
    print()

# ---- numerical error of TF32 operands on the real filter -------------------------------------------------------------
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from sg3_b200 import networks  # noqa: E402
import torch  # noqa: E402

torch.manual_seed(0)
G = networks.Generator(z_dim=512, c_dim=0, w_dim=512, img_resolution=1024, img_channels=3, channel_base=65536, channel_max=1024,
                       conv_kernel=1, use_radial_filters=True)
L = getattr(G.synthesis, G.synthesis.layer_names[11])
F = L.down_filter.numpy().astype(np.float32)
rng = np.random.RandomState(0)
v = rng.randn(6, 300, 300).astype(np.float32) * 2
v = np.clip(np.maximum(v, 0.2 * v) * np.sqrt(2), -256, 256).astype(np.float32)      # what the down filter sees


def tf32(a):
    return (a.view(np.uint32) & np.uint32(0xffffe000)).view(np.float32)              # the tensor core reads the top 19 bits


def down(vv, ff, acc):
    oh, ow = (vv.shape[1] - 12) // 2 + 1, (vv.shape[2] - 12) // 2 + 1
    y = np.zeros((vv.shape[0], oh, ow), acc)
    for a in range(12):
        for b in range(12):
            y += (vv[:, a:a + 2 * oh:2, b:b + 2 * ow:2].astype(acc) * acc(ff[a, b])).astype(acc)
    return y


ref = down(v.astype(np.float64), F.astype(np.float64), np.float64)
scale = np.abs(ref).max()
e32 = np.abs(down(v, F, np.float32) - ref).max() / scale
vh, fh = tf32(v), tf32(F)
e1 = np.abs(down(vh, fh, np.float32) - ref).max() / scale
vl, fl = tf32(v - vh), tf32(F - fh)
y3 = down(vh, fh, np.float32) + down(vh, fl, np.float32) + down(vl, fh, np.float32)
e3 = np.abs(y3 - ref).max() / scale
print('Numerical error of the L11 down pass (radial 12x12 filter of the config-R generator, lrelu-shaped activations), max |err| / max |ref|:')
print()
print(f'* fp32 FMA (the SIMT kernel): {e32:.1e}')
print(f'* TF32 operands (truncated to 10 mantissa bits as the tensor core reads them), fp32 accumulate: {e1:.1e}')
print(f'* 3xTF32 (hi*hi + hi*lo + lo*hi): {e3:.1e}')
