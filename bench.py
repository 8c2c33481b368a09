"""bench.py -- StyleGAN3-R 1024^2 generator-forward throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference] [--config R|T|restyle|pti]

--config R (default) is the metric BASELINE.json is quoted on (configs[2]); T = configs[1]; restyle = configs[3] (ReStyle-pSp
iterative inversion, 5 steps, batch 8); pti = configs[4] (PTI fine-tuning over 64 frames, batch 4 per GPU, data-parallel with
one NCCL gradient all-reduce per step).

One process per GPU (the driver launches N>1 with torch.distributed.run); the batch is sharded
across ranks with no data-path collective (weak scaling, fixed per-GPU batch).  A "step" is one
`G.synthesis(ws)` forward of the per-GPU batch on synthetic latents with random-init weights.
Rank 0 prints ONE JSON line.  Keys:
  value          images/s, whole job, inputs resident in HBM (CUDA events, max over ranks)
  e2e            images/s through the public API with HOST buffers: pinned ws -> H2D, forward, D2H of the images
  roofline       filtered_lrelu (the dominant kernel): algorithmic bytes / CUDA-event time vs measured HBM peak,
                 plus the FP32-pipe fraction that actually binds it
  cpu_baseline   the CPU oracle (port of the reference's impl='ref' path) timed on this box's cores on a bounded sample
`--impl reference` times the reference's own CPU algorithm (oracle port; the Python reference tree cannot
travel to the GPU box) with all host threads; a step is ONE WHOLE image of the same workload (no extrapolation).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

R1024 = dict(z_dim=512, c_dim=0, w_dim=512, img_resolution=1024, img_channels=3,
             channel_base=65536, channel_max=1024, conv_kernel=1, use_radial_filters=True)
T1024 = dict(z_dim=512, c_dim=0, w_dim=512, img_resolution=1024, img_channels=3,
             channel_base=32768, channel_max=512, conv_kernel=3, use_radial_filters=False)
METRIC = 'StyleGAN3-R 1024^2 G-forward images/sec'
CFG_NAME = 'R'          # --config T switches R1024 / METRIC to the config-T generator (BASELINE.json configs[1])


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=32, help='images per GPU per step')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--math', default='tf32', choices=['tf32', 'fp32', 'fp32x3'],
                    help='modulated_conv2d contraction: TF32 tensor cores (the reference\'s cuDNN default), exact FP32 SIMT, or '
                         '3xTF32 tensor cores (fp32-accurate)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--config', default='R', choices=['R', 'T', 'restyle', 'pti'],
                    help='R: StyleGAN3-R 1024^2 forward (the metric BASELINE.json is quoted on, default); T: StyleGAN3-T 1024^2 '
                         'forward (configs[1]); restyle: ReStyle-pSp inversion, 5 steps, batch 8 (configs[3]); pti: PTI fine-tuning, '
                         '64 frames, batch 4 per GPU, NCCL gradient all-reduce (configs[4])')
    ap.add_argument('--frames', type=int, default=64, help='--config pti: global number of frames')
    ap.add_argument('--generator', default='R', choices=['R', 'T'],
                    help='--config pti: the generator being fine-tuned (R = the FFHQ model of configs[4]; T = the 3x3-conv "landscape" '
                         'model, models/stylegan3/model.py:30-40)')
    ap.add_argument('--eager', action='store_true', help='--config restyle: no CUDA-graph replay of the synthesis calls')
    args = ap.parse_args()
    if args.config == 'T':
        global METRIC, CFG_NAME
        R1024.clear(); R1024.update(T1024)
        METRIC = 'StyleGAN3-T 1024^2 G-forward images/sec'
        CFG_NAME = 'T'
    return args


# --------------------------------------------------------------------------------------------------
# Algorithmic work of the workload: per image, fp32.

def layer_work(specs, esize=4):
    """Per-layer algorithmic bytes of filtered_lrelu and FLOPs of the conv, per image."""
    rows = []
    for sp in specs:
        k = sp['conv_kernel']
        conv_hw = sp['in_size'] + k - 1
        fl_bytes = esize * sp['out_channels'] * (conv_hw ** 2 + sp['out_size'] ** 2) + 4 * sp['out_channels']
        conv_flops = 2 * sp['out_channels'] * sp['in_channels'] * k * k * conv_hw ** 2
        # polyphase FMAs of the fused op: H-up on in rows, V-up on up rows, down filter per output
        upw = conv_hw * sp['up'] + sp['padding'][0] + sp['padding'][1] - (sp['up_taps'] - 1)
        fut = max(sp['up_taps'] // sp['up'], 1)
        fd = sp['down_filter']
        dn = 1 if fd is None else (fd.shape[0] * fd.shape[1] if fd.ndim == 2 else 2 * fd.shape[0])
        fma = sp['out_channels'] * (conv_hw * upw * fut + upw * upw * fut + sp['out_size'] ** 2 * dn) if sp['up_taps'] > 1 else 0
        rows.append(dict(name=sp['name'], flrelu_bytes=fl_bytes, conv_flops=conv_flops, flrelu_fma=fma))
    return rows


# --------------------------------------------------------------------------------------------------
# CPU arm: oracle port of the reference's impl='ref' path, bounded sample.

def cpu_image(threads=None):
    """ONE whole image of the workload's synthesis forward on the CPU oracle (the reference's impl='ref' algorithm), all
    layers; returns the wall time in seconds and the thread count.  No extrapolation of any kind."""
    import torch
    from oracle import sg3_oracle as orc
    import sg3_b200  # noqa: F401  (only for the random-init weights; no kernel is launched here)
    from sg3_b200 import networks
    if not threads:
        # all host cores of this process: torchrun exports OMP_NUM_THREADS=1, which would make the CPU arm 16x slower than it is
        threads = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    orc.set_num_threads(threads)
    cores = orc.num_threads()
    st = _CPU_STATE
    if 'net' not in st:
        torch.manual_seed(0)
        G = networks.Generator(**R1024).eval().requires_grad_(False)
        z = torch.randn(1, 512, generator=torch.Generator().manual_seed(1))
        # mapping network on CPU in plain torch (negligible work; bias_act has no CPU path by design)
        x = z * (z.square().mean(1, keepdim=True) + 1e-8).rsqrt()
        for i in range(2):
            fc = getattr(G.mapping, f'fc{i}')
            x = torch.nn.functional.leaky_relu(x @ (fc.weight * fc.weight_gain).t() + fc.bias * fc.bias_gain, 0.2) * np.sqrt(2)
        st['ws'] = x.unsqueeze(1).repeat(1, G.num_ws, 1).numpy()
        state = {k: v.numpy() for k, v in G.synthesis.state_dict().items()}
        cfg = {k: v for k, v in R1024.items() if k not in ('z_dim', 'c_dim', 'w_dim', 'img_resolution', 'img_channels')}
        st['net'] = orc.SynthesisOracle(state, img_resolution=1024, w_dim=512, **cfg)
    t0 = time.perf_counter()
    img = st['net'].forward(st['ws'])
    dt = time.perf_counter() - t0
    assert img.shape == (1, 3, 1024, 1024)
    return dt, cores


_CPU_STATE = {}


def cpu_baseline_entry(dt, cores):
    return dict(value=1.0 / dt, unit='images/s', cores=cores, kind='port',
                sample=f'1 whole image of StyleGAN3-{CFG_NAME} 1024^2 synthesis forward (all 15 layers) on the CPU oracle = port of the '
                       f"reference's impl='ref' path, {dt:.1f} s on {cores} threads; no extrapolation")


def workload_config(world, B, math):
    return dict(workload=f'StyleGAN3-{CFG_NAME} 1024^2 synthesis forward (BASELINE.json configs[{2 if CFG_NAME == "R" else 1}]), random-init seed 0, '
                         'force_fp32, noise_mode=const', per_gpu_batch=B, global_batch=world * B,
                parallelism=f'batch-sharded x{world}, no collective', conv_math=math,
                l2='activations of every layer exceed the 126 MB L2 (inputs larger than L2, no flush needed)')


def run_reference(args):
    """Reference arm: the reference's CPU algorithm on this box's host cores.  One step = one whole image (the per-GPU batch of
    the GPU arm is 32 images of the same kind); `value` = images per second of the timed steps."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    if args.config in ('restyle', 'pti'):
        print(json.dumps(dict(impl='reference', unavailable=f'--config {args.config}: the CPU oracle restates the synthesis hot path only '
                                                               '(no encoder / LPIPS / whole-network backward); use --config R or T')))
        return
    # Bounded run: a whole image takes ~14 s on 16 host threads, so the driver's --steps 20 --warmup 5 would be six minutes of CPU
    # time.  At most ONE warm-up image (a CPU has no clocks / caches to warm beyond that) and timed images until K of them are done
    # or REF_BUDGET_S seconds have passed (at least 2); `steps` reports the images actually timed, `steps_requested` the K asked for.
    REF_BUDGET_S = 150.0
    times, cores = [], 1
    t_begin = time.perf_counter()
    if args.warmup > 0:
        cpu_image()
    while len(times) < args.steps and (len(times) < min(2, args.steps) or time.perf_counter() - t_begin + (times[-1] if times else 0) < REF_BUDGET_S):
        dt, cores = cpu_image()
        times.append(dt)
    steps_requested, args.steps = args.steps, len(times)
    step_s = float(np.mean(times))
    v = 1.0 / step_s
    info = cpu_baseline_entry(step_s, cores)
    cfg = workload_config(args.gpus, args.batch, args.math)
    cfg['reference_sample'] = ('one whole image per step on the host CPU (same generator, same seeds); '
                               f'{len(times)} of the {steps_requested} requested steps timed inside a {REF_BUDGET_S:.0f} s budget, 1 warm-up image')
    out = dict(metric=METRIC, value=v, unit='images/s', impl='reference', n_gpus=args.gpus, steps=args.steps,
               steps_requested=steps_requested, warmup=min(args.warmup, 1), ms_per_step=1e3 * step_s, higher_is_better=True, scaling='weak',
               vs_baseline=None, dtype='f32', data='synthetic', config=cfg,
               cpu_baseline=info, e2e=dict(value=v, unit='images/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(out))


# --------------------------------------------------------------------------------------------------

class ClockSampler:
    """nvidia-smi clocks and throttle reasons sampled while the timed region runs."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-i', str(self.index), '-lms', '200'], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.25)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace('.', '').isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace('.', '').isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
                    if val.lower().startswith('active'):
                        reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def run_ours(args):
    import torch
    import torch.distributed as dist
    import sg3_b200
    from sg3_b200 import capi, filtered_lrelu as fl_mod, modulated_conv, networks, sharding

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

    torch.manual_seed(0)
    G = networks.Generator(**R1024).eval().requires_grad_(False).to(dev)
    B = args.batch
    z_all = torch.randn(world * B, 512, generator=torch.Generator().manual_seed(1))      # the global batch of latents ...
    z = sharding.shard_batch(z_all, rank, world).to(dev)                                 # ... sharded contiguously, no collective
    with torch.no_grad():
        ws = G.mapping(z, None).contiguous()
    ws_host = ws.cpu().pin_memory()
    img_host = [torch.empty([B, 3, 1024, 1024], dtype=torch.float32).pin_memory() for _ in range(2)]
    pipe = sharding.HostPipeline(G, dev)

    # per-call CUDA-event timing of filtered_lrelu (the dominant kernel) inside the timed region
    fl_events = []
    orig_fl = fl_mod.filtered_lrelu
    record = {'on': False}

    def timed_fl(x, *a, **k):
        if not record['on']:
            return orig_fl(x, *a, **k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        y = orig_fl(x, *a, **k)
        e1.record()
        esz = x.element_size()
        fl_events.append((e0, e1, esz * (x.numel() + y.numel()) + 4 * x.shape[1]))
        return y
    fl_mod.filtered_lrelu = timed_fl

    # same for modulated_conv2d (weight prologue + contraction), the second-largest item: tensor-pipe evidence
    cv_events = []
    orig_cv = networks.modulated_conv2d

    def timed_cv(*a, **k):
        if not record['on']:
            return orig_cv(*a, **k)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        y = orig_cv(*a, **k)
        c1.record()
        cv_events.append((c0, c1))
        return y
    networks.modulated_conv2d = timed_cv

    def step_resident():
        with torch.no_grad():
            return G.synthesis(ws, noise_mode='const', force_fp32=True)

    e2e_i = [0]

    def step_e2e():
        # public host-to-host call: pinned latents -> H2D -> forward -> D2H of the images (overlapped with the next step)
        pipe.submit(ws_host, img_host[e2e_i[0] & 1])
        e2e_i[0] += 1

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        return sharding.max_over_ranks(ms, device=dev)

    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()

    # ---- timed: resident inputs ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = capi.lib().sg3_launch_count()
    record['on'] = True
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    barrier()
    record['on'] = False
    launches = capi.lib().sg3_launch_count() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    fl_ms = sum(a.elapsed_time(b) for a, b, _ in fl_events)
    fl_bytes = sum(nb for _, _, nb in fl_events)
    fl_mod.filtered_lrelu = orig_fl
    networks.modulated_conv2d = orig_cv
    cv_ms = sum(a.elapsed_time(b) for a, b in cv_events)

    # ---- timed: end to end with host buffers ----
    for _ in range(2):
        step_e2e()
    pipe.finish()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    pipe.finish()                       # every image of every step is in host memory before the clock stops
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)

    # ---- extra (not the headline): the same step through networks.PipelinedSynthesis (convs and stencils of two micro-batches
    #      co-scheduled on the SMs, profiles/r02_overlap.md); a few steps, CUDA events, max over ranks ----
    pipelined = None
    try:
        pipe2 = networks.PipelinedSynthesis(G.synthesis, micro_batches=2)
        out_img = torch.empty([B, 3, 1024, 1024], dtype=torch.float32, device=dev)
        for _ in range(2):
            pipe2(ws, out=out_img, noise_mode='const', force_fp32=True)
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_pipe = min(args.steps, 5)
        p0.record()
        for _ in range(n_pipe):
            pipe2(ws, out=out_img, noise_mode='const', force_fp32=True)
        p1.record()
        barrier()
        pipe_ms = max_over_ranks(p0.elapsed_time(p1))
        pipelined = dict(value=world * B * n_pipe / (pipe_ms * 1e-3), unit='images/s', ms_per_step=pipe_ms / n_pipe, steps=n_pipe, micro_batches=2,
                         note='opt-in networks.PipelinedSynthesis; the headline `value` is the plain forward')
        del out_img
    except Exception as e:            # never lose the bench line over the secondary figure
        pipelined = dict(error=repr(e))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    # layer geometry for the algorithmic-work table, read off the generator that was just timed
    specs = []
    for lname in G.synthesis.layer_names:
        L = getattr(G.synthesis, lname)
        specs.append(dict(name=lname, conv_kernel=int(L.conv_kernel), in_size=int(L.in_size[0]), out_size=int(L.out_size[0]),
                          in_channels=int(L.in_channels), out_channels=int(L.out_channels), up=int(L.up_factor),
                          up_taps=int(L.up_taps), padding=list(L.padding),
                          down_filter=None if L.down_filter is None else L.down_filter.detach().cpu().numpy()))
    work = layer_work(specs)
    fma_per_img = sum(r['flrelu_fma'] for r in work)
    n_calls = len(fl_events)
    achieved = fl_bytes / (fl_ms * 1e-3) / 1e9 if fl_ms > 0 else 0.0
    fp32_peak_tfma = 148 * 128 * 1.965e9 / 1e12          # 37.2 TFMA/s at max SM clock
    fma_rate = fma_per_img * B * args.steps / (fl_ms * 1e-3) / 1e12 if fl_ms > 0 else 0.0

    # DRAM traffic of the filtered_lrelu launches: NOT measured in this run (a number taken under a profiler is not a bench
    # value) -- a static figure from the committed ncu capture of the same kernels (dram__bytes_read.sum + dram__bytes_write.sum
    # over the 15 launches of one forward), scaled from the profiled batch to this step's batch; compare with `algorithmic_bytes`
    traffic, traffic_note = None, 'static: no committed ncu traffic capture for this config'
    for name in ('r02_flrelu_traffic.json', 'r01_flrelu_traffic.json'):
        try:
            tr = json.load(open(os.path.join(ROOT, 'profiles', name)))
            if tr.get('config') == CFG_NAME:
                traffic = tr['dram_bytes_per_image'] * B
                traffic_note = (f'static (not measured in this run): ncu dram read+write per image from profiles/{name} '
                                f'(batch {tr.get("batch", 2)} capture) x per-GPU batch')
                break
        except Exception:
            pass

    # modulated_conv2d: FLOPs 2*N*O*I*k^2*(H+k-1)^2 and minimum bytes 4*N*(I*H^2 + O*(H+k-1)^2) per layer; the TF32
    # tensor peak is taken as half the measured dense bf16 cuBLAS rate (tcgen05 kind::tf32 runs at half the kind::f16 rate)
    conv = None
    try:
        if cv_ms > 0:
            flops = sum(r['conv_flops'] for r in work) * B * args.steps
            cbytes = sum(4 * (sp['in_channels'] * sp['in_size'] ** 2 + sp['out_channels'] * (sp['in_size'] + sp['conv_kernel'] - 1) ** 2)
                         for sp in specs) * B * args.steps
            tf32_peak = float(peaks.get('bf16_tflops_sustained', peaks.get('bf16_tflops', 2250.0 * 0.62))) / 2
            conv = dict(kernel='modulated_conv2d = weight prologue + tcgen05 TF32 contraction (15 calls per step, CUDA events)',
                        ms_per_step=cv_ms / args.steps, achieved_tflops=flops / (cv_ms * 1e-3) / 1e12, tf32_peak_tflops=tf32_peak,
                        tensor_frac=flops / (cv_ms * 1e-3) / 1e12 / tf32_peak,
                        achieved_gbs=cbytes / (cv_ms * 1e-3) / 1e9, hbm_frac=cbytes / (cv_ms * 1e-3) / 1e9 / hbm_peak,
                        bound='hbm' if CFG_NAME == 'R' else 'tensor',
                        note='config R: 1x1 convs at 59 FLOP/B sit under the HBM roof on tensor cores; config T: 3x3 convs are tensor bound (DESIGN.md 4.3/4.4)')
    except Exception as e:            # never lose the bench line over the secondary table
        conv = dict(error=repr(e))

    value = world * B * args.steps / (ms_total * 1e-3)
    out = dict(
        metric=METRIC, value=value, unit='images/s', n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
        ms_per_step=ms_total / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
        dtype={'fp32': 'f32', 'tf32': 'f32 (tf32 tensor-core conv)', 'fp32x3': 'f32 (3xTF32 tensor-core conv, fp32-accurate)'}[args.math],
        data='synthetic',
        config=workload_config(world, B, args.math),
        build=capi.lib().sg3_build_info().decode(),
        clocks=clocks,
        e2e=dict(value=world * B * args.steps / (e2e_ms * 1e-3), unit='images/s',
                 h2d_bytes_per_step=int(ws_host.numel() * 4), d2h_bytes_per_step=int(img_host[0].numel() * 4)),
        gpu_launches=int(launches),
        conv=conv,
        pipelined=pipelined,
        # `bound` names the roof `achieved` / `peak` / `frac` are quoted against (the contract's HBM roof); the roof that actually
        # binds this kernel with fp32 SIMT math is the FP32 pipe: `binds` + the flat fp32_pipe_* keys
        roofline=dict(bound='hbm', achieved=achieved, peak=hbm_peak, unit='GB/s', frac=achieved / hbm_peak,
                      traffic=traffic, algorithmic_bytes=fl_bytes / args.steps, traffic_note=traffic_note,
                      kernel='filtered_lrelu (15 calls per step, all timed with CUDA events)',
                      launches_timed=n_calls, ms_per_step=fl_ms / args.steps,
                      peak_source='MEASURED_PEAKS.json hbm_gbs (of measured)' if peaks else 'fallback 6650 GB/s (of fallback)',
                      binds='fp32_pipe', fp32_pipe_frac=fma_rate / fp32_peak_tfma, fp32_pipe_achieved_tfma=fma_rate,
                      fp32_pipe_peak_tfma=fp32_peak_tfma,
                      fp32_pipe_note='nominal polyphase FMAs of the fused op (dense 12x12 down filter counted as 144 MACs per output) / '
                                     'CUDA-event time vs 148 SMs x 128 lanes x 1.965 GHz'),
    )
    # cpu_baseline: timed on rank 0 at N = 1 only (the host cores are shared by all ranks of a multi-GPU run)
    out['cpu_baseline'] = cpu_baseline_entry(*cpu_image()) if (world == 1 and not args.no_cpu_baseline) else None
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference(args)
    elif args.config == 'restyle':
        from examples import bench_workloads
        bench_workloads.run_restyle(args, ClockSampler)
    elif args.config == 'pti':
        from examples import bench_workloads
        bench_workloads.run_pti(args, ClockSampler)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
