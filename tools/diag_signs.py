import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sg3_b200
from sg3_b200 import filtered_lrelu as fl
from oracle import sg3_oracle as orc
fl._quiet_fallback = True
def unpack(s, w):
    s = s.cpu().numpy().astype(np.uint8)
    return np.stack([(s >> (2 * k)) & 3 for k in range(4)], axis=-1).reshape(*s.shape[:3], -1)[..., :w]
for (up, radial, N, C, size, pad) in [(2, True, 1, 2, 84, [11, 10, 11, 10]), (4, True, 1, 2, 52, [-2, -5, -2, -5]), (2, False, 2, 3, 84, [11,10,11,10]), (2, False, 1, 2, 84, [-9,-10,-9,-10])]:
    fu = orc.design_lowpass_filter(6 * up, 11.3, 26.0, 128 if up == 2 else 256)
    fd = orc.design_lowpass_filter(12, 16.0, 36.0, 128, radial=radial)
    x = torch.randn(N, C, size, size, device='cuda') * 3
    b = torch.randn(C, device='cuda')
    cfg = (up, 2) + tuple(pad) + (float(np.sqrt(2)), 0.2, 256.0, False)
    fut, fdt = torch.from_numpy(fu).cuda(), torch.from_numpy(fd).cuda()
    y1, s1 = fl._fused(x, fut, fdt, b, None, 0, 0, cfg, True)
    y2, s2 = fl._generic(x, fut, fdt, b, None, 0, 0, cfg, True)
    sw = 2 * y1.shape[3] + 10
    sh = s1.shape[2]
    u1, u2 = unpack(s1, sw), unpack(s2[:, :, :sh], sw)
    d = (u1 != u2)
    print('case', up, radial, size, 'y diff', float((y1 - y2).abs().max()), 'sign mismatches', int(d.sum()), 'of', d.size, 'shape', tuple(s1.shape), tuple(s2.shape))
    if d.sum():
        idx = np.argwhere(d)
        print('  rows', np.unique(idx[:, 2])[:20], 'cols', np.unique(idx[:, 3])[:40])
    # backward both ways
    dy = torch.randn_like(y1)
    fu_w, fd_w = fu.shape[-1], fd.shape[-1]
    xw = size; yw = y1.shape[3]
    adj = (2, up, (fu_w - 1) + (fd_w - 1) - pad[0], xw * up - yw * 2 + pad[0] - (up - 1), (fu_w - 1) + (fd_w - 1) - pad[2], xw * up - yw * 2 + pad[2] - (up - 1),
           float(np.sqrt(2)) * up ** 2 / 4, 0.2, float('inf'), True)
    sx, sy = -(fu_w - 1) + pad[0], -(fu_w - 1) + pad[2]
    dxa, _ = fl._generic(dy, fdt, fut, None, s1, sx, sy, adj, False)
    dxb, _ = fl._generic(dy, fdt, fut, None, s2, sx, sy, adj, False)
    print('  dx(generic bwd) with fused signs vs generic signs: max diff', float((dxa - dxb).abs().max()), 'scale', float(dxb.abs().max()))
    r = fl._fused(dy, fdt, fut, None, s1, sx, sy, adj, False)
    if r is not None:
        print('  dx fused-read vs generic-read', float((r[0] - dxa).abs().max()))
print('--- vs oracle ---')
up, radial, N, C, size, pad = 2, False, 1, 2, 84, [-9, -10, -9, -10]
fu = orc.design_lowpass_filter(12, 11.3, 26.0, 128); fd = orc.design_lowpass_filter(12, 16.0, 36.0, 128)
rng = np.random.RandomState(5)
xn = (rng.randn(N, C, size, size) * 3).astype(np.float32); bn = rng.randn(C).astype(np.float32)
yo, so = orc.filtered_lrelu(xn, fu, fd, bn, up=2, down=2, padding=pad, gain=np.sqrt(2), slope=0.2, clamp=256, return_signs=True)
x = torch.from_numpy(xn).cuda(); b = torch.from_numpy(bn).cuda()
cfg = (2, 2) + tuple(pad) + (float(np.sqrt(2)), 0.2, 256.0, False)
fut, fdt = torch.from_numpy(fu).cuda(), torch.from_numpy(fd).cuda()
y1, s1 = fl._fused(x, fut, fdt, b, None, 0, 0, cfg, True)
y2, s2 = fl._generic(x, fut, fdt, b, None, 0, 0, cfg, True)
sw = 138
uo = unpack(torch.from_numpy(so), sw)
for nm, s in (('fused', s1), ('generic', s2)):
    u = unpack(s[:, :, :so.shape[2]], sw)
    d = u != uo
    print(nm, 'mismatch vs oracle', int(d.sum()), 'rows', np.unique(np.argwhere(d)[:, 2])[:12], 'cols', np.unique(np.argwhere(d)[:, 3])[:12])
    if d.sum():
        i = np.argwhere(d)[0]; print('   first', i, 'got', u[tuple(i)], 'oracle', uo[tuple(i)])
