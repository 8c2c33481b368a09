"""bench.py --config restyle | pti: BASELINE.json configs[3] and configs[4] as measured workloads (same JSON contract as the
forward bench: device-timed `value`, host-buffer `e2e`, `roofline` of filtered_lrelu from CUDA events, launch count, clocks).

One process per GPU.  restyle: every rank inverts its own batch of 8 images (no collective, weak scaling).  pti: the 64 frames are
sharded contiguously across ranks, batch 4 per GPU and step, ONE flat fp32 gradient all-reduce over NCCL per step (weak scaling
in frames per step: world x 4), identical Adam step on every rank.
"""
import json
import os
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

R1024 = dict(z_dim=512, c_dim=0, w_dim=512, img_resolution=1024, img_channels=3,
             channel_base=65536, channel_max=1024, conv_kernel=1, use_radial_filters=True)
T1024 = dict(z_dim=512, c_dim=0, w_dim=512, img_resolution=1024, img_channels=3,
             channel_base=32768, channel_max=512, conv_kernel=3, use_radial_filters=False)


def _setup(args):
    import sg3_b200
    from sg3_b200 import capi, filtered_lrelu as fl_mod, modulated_conv, sharding
    rank, world, local = sharding.rank_info()
    assert torch.cuda.is_available(), 'bench.py needs CUDA (no CPU fallback for the product path)'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    capi.lib()
    modulated_conv.set_math(args.math)
    fl_mod._quiet_fallback = True
    return rank, world, local, dev


class _FlreluTimer:
    """CUDA-event timing of every fused filtered_lrelu launch (forward, forward with sign write, backward) inside the timed
    region; bytes = the algorithmic bytes of the call (x + y, + the sign tensor when written / read)."""

    def __init__(self):
        from sg3_b200 import filtered_lrelu as fl_mod
        self.mod, self.orig, self.on, self.events = fl_mod, fl_mod._fused, False, []
        fl_mod._fused = self._call

    def _call(self, x, fu, fd, b, si, sx, sy, cfg, write_signs, **kw):
        if not self.on:
            return self.orig(x, fu, fd, b, si, sx, sy, cfg, write_signs, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = self.orig(x, fu, fd, b, si, sx, sy, cfg, write_signs, **kw)
        e1.record()
        if res is not None:
            y, so = res
            nb = x.element_size() * (x.numel() + y.numel())
            nb += so.numel() if so is not None else (si.numel() if si is not None else 0)
            self.events.append((e0, e1, nb))
        return res

    def close(self):
        self.mod._fused = self.orig
        ms = sum(a.elapsed_time(b) for a, b, _ in self.events)
        return ms, sum(nb for _, _, nb in self.events), len(self.events)


def _barrier(world):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        return {}


def _roofline(fl_ms, fl_bytes, n_calls, steps, what):
    peaks = _peaks()
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    achieved = fl_bytes / (fl_ms * 1e-3) / 1e9 if fl_ms > 0 else 0.0
    return dict(bound='hbm', achieved=achieved, peak=hbm_peak, unit='GB/s', frac=achieved / hbm_peak, traffic=None,
                algorithmic_bytes=fl_bytes / steps, kernel=what, launches_timed=n_calls, ms_per_step=fl_ms / steps,
                peak_source='MEASURED_PEAKS.json hbm_gbs (of measured)' if peaks else 'fallback 6650 GB/s (of fallback)',
                binds='fp32_pipe', traffic_note='no ncu traffic capture for this workload')


def run_restyle(args, ClockSampler):
    from sg3_b200 import capi, networks, sharding
    from examples import workloads
    rank, world, local, dev = _setup(args)
    B = 8 if args.batch == 32 else args.batch               # configs[3]: batch 8 (32 is the forward bench's default)
    n_iters = 5
    torch.manual_seed(0)
    G = networks.Generator(**R1024).eval().requires_grad_(False).to(dev)
    torch.manual_seed(1)
    enc = workloads.RestyleEncoder(n_styles=G.num_ws, input_nc=6).eval().requires_grad_(False).to(dev).to(memory_format=torch.channels_last)
    net = workloads.PSP(enc, G, use_graph=not args.eager)
    gen = torch.Generator().manual_seed(100 + rank)
    inputs_host = (torch.rand(B, 3, 256, 256, generator=gen) * 2 - 1).pin_memory()
    inputs = inputs_host.to(dev)
    lm = workloads.random_landmarks_transforms(B, gen, dev)
    with torch.no_grad():
        avg_image = net.face_pool(G.synthesis(net.latent_avg.repeat(G.num_ws, 1).unsqueeze(0), noise_mode='const', force_fp32=True))[0]
    out_host = torch.empty(B, 3, 256, 256).pin_memory()
    lat_host = torch.empty(B, G.num_ws, 512).pin_memory()

    def step_resident():
        return workloads.run_on_batch(inputs, net, avg_image, n_iters=n_iters, landmarks_transform=lm)

    def step_e2e():
        x = inputs_host.to(dev, non_blocking=True)
        y, lat = workloads.run_on_batch(x, net, avg_image, n_iters=n_iters, landmarks_transform=lm)
        out_host.copy_(y, non_blocking=True)
        lat_host.copy_(lat, non_blocking=True)
        torch.cuda.current_stream().synchronize()            # the caller reads the inversion result

    for _ in range(max(args.warmup, 3)):
        step_resident()
    _barrier(world)
    # the graph replays hide the per-call Python wrapper, so the per-kernel event timing runs on an eager twin of the step
    timer = _FlreluTimer()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = capi.lib().sg3_launch_count()
    _barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    _barrier(world)
    launches_host = capi.lib().sg3_launch_count() - launches0
    ms_total = sharding.max_over_ranks(e0.elapsed_time(e1), device=dev)
    clocks = sampler.stop() if rank == 0 else None
    # kernel time of filtered_lrelu: one eager (ungraphed) inversion with events around every launch
    eager = workloads.PSP(enc, G, use_graph=False)
    # one untimed eager inversion first: the graphs own a private memory pool, so the first eager pass allocates with cudaMalloc,
    # and an allocation between the two events of a launch would be counted as kernel time
    workloads.run_on_batch(inputs, eager, avg_image, n_iters=n_iters, landmarks_transform=lm)
    torch.cuda.synchronize()
    launches1 = capi.lib().sg3_launch_count()
    timer.on = True
    workloads.run_on_batch(inputs, eager, avg_image, n_iters=n_iters, landmarks_transform=lm)
    torch.cuda.synchronize()
    timer.on = False
    fl_ms, fl_bytes, n_calls = timer.close()
    launches_per_step = int(capi.lib().sg3_launch_count() - launches1)

    for _ in range(2):
        step_e2e()
    _barrier(world)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    _barrier(world)
    e2e_ms = sharding.max_over_ranks((time.perf_counter() - t0) * 1e3, device=dev)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    out = dict(
        metric='ReStyle-pSp + StyleGAN3-R 1024^2 iterative inversion, images inverted/sec', value=world * B * args.steps / (ms_total * 1e-3),
        unit='images/s', n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=ms_total / args.steps,
        higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32 (tf32 tensor-core conv)', data='synthetic',
        config=dict(workload='BASELINE.json configs[3]: ReStyle-pSp encoder (IR-SE50 backbone, random-init) + StyleGAN3-R 1024^2 generator, '
                             f'{n_iters} refinement steps per batch, two synthesis calls per step (identity + per-sample landmarks transform), '
                             'face_pool 1024 -> 256', per_gpu_batch=B, global_batch=world * B,
                    parallelism=f'batch-sharded x{world}, no collective', conv_math=args.math,
                    synthesis='eager' if args.eager else 'CUDA-graph replay (networks.GraphedSynthesis), transforms updated in place',
                    l2='activations of the 1024^2 layers exceed the 126 MB L2'),
        build=capi.lib().sg3_build_info().decode(), clocks=clocks,
        e2e=dict(value=world * B * args.steps / (e2e_ms * 1e-3), unit='images/s', h2d_bytes_per_step=int(inputs_host.numel() * 4),
                 d2h_bytes_per_step=int((out_host.numel() + lat_host.numel()) * 4)),
        gpu_launches=int(launches_per_step * args.steps),
        gpu_launches_note='sg3 kernels per inversion step x steps (counted on an eager step; the timed steps replay them from CUDA graphs)',
        roofline=_roofline(fl_ms, fl_bytes, n_calls, 1, f'filtered_lrelu ({n_calls} calls per inversion step, CUDA events on an eager step)'),
        cpu_baseline=None)
    out['roofline']['share_of_step'] = fl_ms / (ms_total / args.steps)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_pti(args, ClockSampler):
    from sg3_b200 import capi, networks, sharding
    from examples import workloads
    rank, world, local, dev = _setup(args)
    B = 4 if args.batch == 32 else args.batch               # run_pti_video.py:43 batch_size = 4
    frames_total = args.frames
    torch.manual_seed(100 + rank)                            # deliberately different init per rank: the broadcast must fix it
    gen_name = getattr(args, 'generator', 'R')
    G = networks.Generator(**(T1024 if gen_name == 'T' else R1024)).to(dev)
    sharding.broadcast_parameters(G)
    gen = torch.Generator().manual_seed(7)
    z = torch.randn(frames_total, 512, generator=gen)
    b0, b1 = sharding.shard_range(frames_total, rank, world)
    # frames are generated per shard (the global set would be 805 MB of host memory per rank for nothing)
    fgen = torch.Generator().manual_seed(1000 + b0)
    frames_host = (torch.rand(b1 - b0, 3, 1024, 1024, generator=fgen) * 2 - 1).pin_memory()
    with torch.no_grad():
        ws = G.mapping(z[b0:b1].to(dev), None).contiguous()
    frames = frames_host.to(dev)
    trainer = workloads.PTITrainer(G, frames, ws, batch=B, world=world, use_graph=not args.eager)
    ws_host = ws.cpu().pin_memory()

    def step_resident():
        return trainer.step()

    def step_e2e():
        # frames and latents of the step come from pinned host memory; the loss value goes back (the reference prints it)
        idx = trainer._next_batch()
        if idx[-1] == idx[0] + B - 1:
            tgt = frames_host[idx[0]:idx[0] + B]
        else:
            tgt = torch.stack([frames_host[i] for i in idx])
        trainer.load_batch(tgt, ws_host[idx])
        return float(trainer.run_loaded())                   # D2H of the loss: synchronises the step

    for _ in range(max(args.warmup, 3)):
        step_resident()
    _barrier(world)
    timer = _FlreluTimer()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    timer.on = args.eager                                    # a graph replay hides the per-call wrapper: see below
    launches0 = capi.lib().sg3_launch_count()
    _barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    losses = []
    for _ in range(args.steps):
        losses.append(step_resident())
    e1.record()
    _barrier(world)
    timer.on = False
    launches = capi.lib().sg3_launch_count() - launches0
    ms_total = sharding.max_over_ranks(e0.elapsed_time(e1), device=dev)
    clocks = sampler.stop() if rank == 0 else None
    fl_steps = args.steps
    if not args.eager:
        # kernel time of filtered_lrelu and the launch count: two eager steps of the same computation with events around every launch
        for _ in range(2):                                   # untimed: the eager allocations (the graphs own a private pool)
            trainer.load_batch(frames[:B], ws[:B])
            trainer._fwd_bwd()
        torch.cuda.synchronize()
        timer.on = True
        l0 = capi.lib().sg3_launch_count()
        fl_steps = 2
        for _ in range(fl_steps):
            trainer.load_batch(frames[:B], ws[:B])
            trainer._fwd_bwd()
        torch.cuda.synchronize()
        timer.on = False
        launches = (capi.lib().sg3_launch_count() - l0) // fl_steps * args.steps
    fl_ms, fl_bytes, n_calls = timer.close()

    for _ in range(2):
        step_e2e()
    _barrier(world)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    _barrier(world)
    e2e_ms = sharding.max_over_ranks((time.perf_counter() - t0) * 1e3, device=dev)
    # every rank must hold the same parameters after the same updates (torch_utils/misc.py:182-193)
    flat = torch.cat([p.detach().flatten() for p in trainer.bucket.params])
    ref = flat.clone()
    if world > 1:
        dist.broadcast(ref, src=0)
    dev_max = sharding.max_over_ranks(float((flat - ref).abs().max()), device=dev)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    losses = [float(l) for l in losses]
    out = dict(
        metric=f'PTI fine-tuning of StyleGAN3-{gen_name} 1024^2 over video frames, frames/sec (forward + backward + Adam)',
        value=world * B * args.steps / (ms_total * 1e-3), unit='frames/s', n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
        ms_per_step=ms_total / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
        dtype='f32 (tf32 tensor-core conv)', data='synthetic',
        config=dict(workload=f'BASELINE.json configs[4]: PTI over {frames_total} synthetic 1024^2 frames, loss = MSE + LPIPS(AlexNet structure, '
                             'random-init), Adam lr 3e-4 on list(G.synthesis.parameters())[3:]', per_gpu_batch=B, global_batch=world * B,
                    frames=frames_total, frames_per_rank=b1 - b0,
                    parallelism=f'data-parallel over frames x{world}, one flat fp32 gradient all-reduce per step '
                                f'({trainer.bucket.numel} elements) over NCCL' if world > 1 else 'single GPU (no collective)',
                    conv_math=args.math, l2='activations of the 1024^2 layers exceed the 126 MB L2'),
        build=capi.lib().sg3_build_info().decode(), clocks=clocks,
        e2e=dict(value=world * B * args.steps / (e2e_ms * 1e-3), unit='frames/s',
                 h2d_bytes_per_step=int(B * 3 * 1024 * 1024 * 4 + B * ws.shape[1] * 512 * 4), d2h_bytes_per_step=4),
        gpu_launches=int(launches),
        roofline=_roofline(fl_ms, fl_bytes, n_calls, fl_steps,
                           f'filtered_lrelu forward with sign write + backward ({n_calls // max(fl_steps, 1)} launches per step, CUDA events'
                           + ('' if args.eager else ' on eager steps; the timed steps replay the same launches from CUDA graphs') + ')'),
        loss_first=losses[0], loss_last=losses[-1], rank_parameter_divergence=dev_max,
        cpu_baseline=None)
    out['roofline']['share_of_step'] = (fl_ms / fl_steps) / (ms_total / args.steps)
    out['config']['step'] = 'eager' if args.eager else 'two CUDA graphs per step (zero + forward + loss + backward | Adam) around the eager NCCL all-reduce'
    print(json.dumps(out))
    assert dev_max == 0.0, f'ranks diverged by {dev_max}'
    if world > 1:
        dist.destroy_process_group()
