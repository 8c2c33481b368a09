"""Data-parallel PTI-style fine-tuning step of the StyleGAN3 generator over frames (BASELINE.json configs[4]).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 examples/pti_dp.py --res 256

Mirrors the reference's single-GPU loop (inversion/video/run_pti_video.py:96-168, run_pti_images.py:111-186):
Adam(lr=3e-4) on list(G.synthesis.parameters())[3:], loss = MSE + a perceptual term, batch of frames per step --
made data-parallel: every rank owns a contiguous shard of the frames, gradients live in one flat fp32 bucket that
is all-reduced once per step over NCCL (sg3_b200.sharding.FlatGradBucket), identical Adam step on every rank.
Synthetic frames and latents, random-init weights (no datasets or checkpoints in this environment).
"""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sg3_b200  # noqa: E402
from sg3_b200 import modulated_conv, networks, sharding  # noqa: E402


def perceptual(a, b):
    """Stand-in for LPIPS (no pretrained weights offline): multi-scale L1 of average-pooled images."""
    loss = 0
    for k in (2, 4, 8):
        loss = loss + (torch.nn.functional.avg_pool2d(a, k) - torch.nn.functional.avg_pool2d(b, k)).abs().mean()
    return loss / 3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--res', type=int, default=256)
    ap.add_argument('--frames', type=int, default=8, help='global number of frames per step')
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--math', default='tf32')
    args = ap.parse_args()

    rank, world, local = sharding.rank_info()
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    sg3_b200.filtered_lrelu._quiet_fallback = True
    modulated_conv.set_math(args.math)

    torch.manual_seed(100 + rank)          # deliberately different init per rank: the broadcast must fix it
    G = networks.Generator(z_dim=512, c_dim=0, w_dim=512, img_resolution=args.res, img_channels=3, **networks.CONFIG_R).to(dev)
    sharding.broadcast_parameters(G)
    params = list(G.synthesis.parameters())[3:]          # skip the Fourier-feature input, as the reference does
    for p in G.parameters():
        p.requires_grad_(False)
    for p in params:
        p.requires_grad_(True)
    bucket = sharding.FlatGradBucket(params)
    opt = torch.optim.Adam(params, lr=3e-4)

    gen = torch.Generator().manual_seed(7)
    z = torch.randn(args.frames, 512, generator=gen)
    target = (torch.rand(args.frames, 3, args.res, args.res, generator=gen) * 2 - 1)
    with torch.no_grad():
        ws_all = G.mapping(z.to(dev), None)
    ws = sharding.shard_batch(ws_all, rank, world)
    tgt = sharding.shard_batch(target, rank, world).to(dev)

    losses = []
    warm = 2                               # untimed: allocator growth (cudaMalloc of GB-sized activations), lazy init
    t0 = None
    for step in range(args.steps + warm):
        if step == warm:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        bucket.zero()
        img = G.synthesis(ws, noise_mode='const', force_fp32=True)
        # per-rank partial of the global mean loss: sum over own frames / global frame count
        loss = (((img - tgt) ** 2).mean(dim=(1, 2, 3)).sum() + 0.5 * perceptual(img, tgt) * img.shape[0]) / args.frames
        (loss * world).backward()          # all_reduce_mean divides by world
        bucket.all_reduce_mean()
        opt.step()
        l = loss.detach().clone()
        if world > 1:
            dist.all_reduce(l)
        losses.append(float(l))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0

    # ranks must hold identical parameters after identical updates (torch_utils/misc.py:182-193 check_ddp_consistency)
    flat = torch.cat([p.detach().flatten() for p in params])
    ref = flat.clone()
    if world > 1:
        dist.broadcast(ref, src=0)
    max_dev = float((flat - ref).abs().max())
    if rank == 0:
        print(f'world={world} res={args.res} frames/step={args.frames} steps={args.steps} '
              f'loss {losses[0]:.6f} -> {losses[-1]:.6f}  {dt / args.steps * 1e3:.1f} ms/step  '
              f'grad bucket {bucket.numel} fp32  launches {sg3_b200.capi.lib().sg3_launch_count()}')
    assert max_dev == 0.0, f'rank {rank} diverged from rank 0 by {max_dev}'
    assert losses[-1] < losses[0], 'loss did not decrease'
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
