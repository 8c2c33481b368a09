"""Does the conv / stencil overlap of `networks.PipelinedSynthesis` happen, and what is it worth?   (run on a B200)

    python tools/overlap_probe.py [--config R|T] [--batch 32] [--steps 5]

Prints (i) whole-forward time: plain `G.synthesis`, pipelined with 2 / 4 micro-batches, and the pipelined schedule WITHOUT the
shared-memory budget (conv CTAs cannot sit next to stencil CTAs: only kernel tails overlap) -- the difference is the co-residency;
(ii) one layer under the microscope: conv alone, stencil alone, both launched together on the two streams;
(iii) parity of the pipelined image with the plain one.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sg3_b200  # noqa: E402
from sg3_b200 import capi, modulated_conv, networks  # noqa: E402


def timeit(fn, steps, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', default='R')
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--layer', default='L11')
    ap.add_argument('--math', default='tf32')
    ap.add_argument('--skip-net', action='store_true')
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    modulated_conv.set_math(args.math)
    cfg = networks.CONFIG_R if args.config == 'R' else networks.CONFIG_T
    torch.manual_seed(0)
    G = networks.Generator(z_dim=512, c_dim=0, w_dim=512, img_resolution=1024, img_channels=3, **cfg).eval().requires_grad_(False).to(dev)
    z = torch.randn(args.batch, 512, generator=torch.Generator().manual_seed(1)).to(dev)
    with torch.no_grad():
        ws = G.mapping(z, None).contiguous()
    kw = dict(noise_mode='const', force_fp32=True)
    B = args.batch

    def plain():
        with torch.no_grad():
            return G.synthesis(ws, **kw)

    t_plain = timeit(plain, args.steps) if not args.skip_net else 0.0
    print(f'config {args.config} batch {B}: plain forward {t_plain:8.2f} ms  {B / max(t_plain, 1e-9) * 1e3:7.1f} images/s', flush=True)
    ref = plain()
    out = torch.empty_like(ref)
    stats0 = torch.cuda.memory_stats()
    for M in (() if args.skip_net else (2, 4, 8)):
        for budget, tag in ((networks.PipelinedSynthesis.CONV_SMEM_BUDGET, 'budget 138 KB (co-resident)'), (0, 'no budget (no co-residency)')):
            if budget == 0 and M != 2:
                continue
            pipe = networks.PipelinedSynthesis(G.synthesis, micro_batches=M, conv_smem_budget=budget)
            t = timeit(lambda: pipe(ws, out=out, **kw), args.steps)
            err = float((out - ref).abs().max() / ref.abs().max())
            print(f'  pipelined M={M} {tag:28s} {t:8.2f} ms  {B / t * 1e3:7.1f} images/s   x{t_plain / t:5.3f}   max rel diff vs plain {err:.2e}', flush=True)
    stats1 = torch.cuda.memory_stats()
    print('  allocator during the pipelined runs: cudaMalloc calls', stats1['num_device_alloc'] - stats0['num_device_alloc'], 'cudaFree calls',
          stats1['num_device_free'] - stats0['num_device_free'], 'retries', stats1['num_alloc_retries'] - stats0['num_alloc_retries'],
          'peak GB', round(torch.cuda.max_memory_allocated() / 1e9, 1), flush=True)
    for budget in (() if args.skip_net else (100 * 1024, 120 * 1024)):
        pipe = networks.PipelinedSynthesis(G.synthesis, micro_batches=2, conv_smem_budget=budget)
        t = timeit(lambda: pipe(ws, out=out, **kw), args.steps)
        print(f'  pipelined M=2 budget {budget // 1024:3d} KB {"":13s} {t:8.2f} ms  {B / t * 1e3:7.1f} images/s   x{t_plain / t:5.3f}', flush=True)

    # ---- one layer: conv alone, stencil alone, together ----
    name = [n for n in G.synthesis.layer_names if n.startswith(args.layer + '_')][0]
    L = getattr(G.synthesis, name)
    half = B // 2
    x = torch.randn([half, L.in_channels, int(L.in_size[1]), int(L.in_size[0])], device=dev)
    w = ws[:half, 1]
    pipe = networks.PipelinedSynthesis(G.synthesis, conv_smem_budget=networks.PipelinedSynthesis.CONV_SMEM_BUDGET)
    conv_s, act_s = pipe._get_streams(dev)
    with torch.no_grad():
        t0 = L.conv_part(x, w, **kw)

        def conv_only():
            L.conv_part(x, w, **kw)

        def act_only():
            L.act_part(t0)

        def both():
            main = torch.cuda.current_stream()
            conv_s.wait_stream(main)
            act_s.wait_stream(main)
            with torch.cuda.stream(act_s):
                L.act_part(t0)
            with torch.cuda.stream(conv_s):
                L.conv_part(x, w, **kw)
            main.wait_stream(conv_s)
            main.wait_stream(act_s)

        tc = timeit(conv_only, 10)
        ta = timeit(act_only, 10)
        tb0 = timeit(both, 10)
        prev = capi.lib().sg3_modconv_set_smem_budget(pipe.budget)
        tcb = timeit(conv_only, 10)
        tb1 = timeit(both, 10)
        capi.lib().sg3_modconv_set_smem_budget(prev)
    print(f'layer {name} at batch {half}: conv {tc:.3f} ms (with budget {tcb:.3f}), stencil {ta:.3f} ms, sum {tc + ta:.3f}; '
          f'together without budget {tb0:.3f} ms, with budget {tb1:.3f} ms', flush=True)

    # ---- who sat where: CTA residency trace (needs a -DSG3_TRACE build, see build.py: SG3_NVCC_EXTRA / SG3_LIB_SUFFIX) ----
    lib = capi.lib()
    if not hasattr(lib, 'sg3_debug_set_trace'):
        return
    import ctypes
    import numpy as np
    lib.sg3_debug_set_trace.argtypes = [ctypes.c_void_p]
    lib.sg3_debug_set_trace.restype = None
    cap = 400000
    for order in ('stencil first', 'conv first'):
        buf = torch.zeros(2 + 3 * cap, dtype=torch.int64, device=dev)
        buf[1] = cap
        prev = lib.sg3_modconv_set_smem_budget(pipe.budget)
        torch.cuda.synchronize()
        lib.sg3_debug_set_trace(buf.data_ptr())
        with torch.no_grad():
            main = torch.cuda.current_stream()
            conv_s.wait_stream(main)
            act_s.wait_stream(main)
            if order == 'stencil first':
                with torch.cuda.stream(act_s):
                    L.act_part(t0)
                with torch.cuda.stream(conv_s):
                    L.conv_part(x, w, **kw)
            else:
                with torch.cuda.stream(conv_s):
                    L.conv_part(x, w, **kw)
                with torch.cuda.stream(act_s):
                    L.act_part(t0)
            main.wait_stream(conv_s)
            main.wait_stream(act_s)
        torch.cuda.synchronize()
        lib.sg3_debug_set_trace(None)
        lib.sg3_modconv_set_smem_budget(prev)
        h = buf.cpu().numpy()
        n = int(min(h[0], cap))
        rec = h[2:2 + 3 * n].reshape(n, 3)
        kind, smid, ident, t = rec[:, 0] >> 32, rec[:, 0] & 0xffffffff, rec[:, 1], rec[:, 2]
        tmin = t.min()
        t = (t - tmin) * 1e-6                                   # ms

        def spans(k0, k1):
            a = {(int(i)): (int(sm), float(tt)) for i, sm, tt in zip(ident[kind == k0], smid[kind == k0], t[kind == k0])}
            b = {(int(i)): float(tt) for i, tt in zip(ident[kind == k1], t[kind == k1])}
            return [(sm, t0_, b[i]) for i, (sm, t0_) in a.items() if i in b]
        st, cv = spans(1, 2), spans(3, 4)
        print(f'trace [{order}]: {len(st)} stencil warps, {len(cv)} conv CTAs')
        if not st or not cv:
            continue
        st_a = np.array(st)
        cv_a = np.array(cv)
        print(f'  stencil: first start {st_a[:, 1].min():.3f} ms, last end {st_a[:, 2].max():.3f} ms, mean warp life {np.mean(st_a[:, 2] - st_a[:, 1]):.3f} ms')
        print(f'  conv   : first start {cv_a[:, 1].min():.3f} ms, median start {np.median(cv_a[:, 1]):.3f}, last start {cv_a[:, 1].max():.3f}, '
              f'last end {cv_a[:, 2].max():.3f} ms, mean CTA life {np.mean(cv_a[:, 2] - cv_a[:, 1]):.3f} ms')
        # stencil warps resident on the conv CTA's SM, averaged over the conv CTA's life
        res = []
        for sm, a, b in cv:
            mine = st_a[st_a[:, 0] == sm]
            ov = np.clip(np.minimum(mine[:, 2], b) - np.maximum(mine[:, 1], a), 0, None).sum()
            res.append(ov / max(b - a, 1e-9))
        res = np.array(res)
        print(f'  stencil warps co-resident with a conv CTA (time average over the CTA life): mean {res.mean():.2f}, min {res.min():.2f}, max {res.max():.2f}   (16 = full stencil occupancy, 12 = three CTAs)')


if __name__ == '__main__':
    main()
