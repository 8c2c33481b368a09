"""StyleGAN3 generator on the sg3_b200 kernels.

Interface mirror of `models/stylegan3/networks_stylegan3.py` of the reference (class names,
constructor arguments, forward signatures, parameter / buffer names and creation order are the
same, so state dicts load either way and `torch.manual_seed(s); Generator(...)` draws identical
random-init weights).  The layer math runs on this package's ops: `modulated_conv2d` (fused weight
prologue + contraction kernels) and the fused `filtered_lrelu`; the mapping network's
fully-connected layers use `bias_act`.

    G = Generator(z_dim=512, c_dim=0, w_dim=512, img_resolution=1024, img_channels=3,
                  channel_base=65536, channel_max=1024, conv_kernel=1, use_radial_filters=True)   # config R
    img = G.synthesis(G.mapping(z, None), noise_mode='const', force_fp32=True)
"""
import numpy as np
import scipy.signal
import scipy.special
import torch

from . import bias_act, filtered_lrelu
from .modulated_conv import modulated_conv2d, _math_mode, tf32_activation_policy

__all__ = ['FullyConnectedLayer', 'MappingNetwork', 'SynthesisInput', 'SynthesisLayer', 'SynthesisNetwork',
           'Generator', 'GraphedSynthesis', 'PipelinedSynthesis', 'modulated_conv2d', 'CONFIG_R', 'CONFIG_T']

# Keyword sets of the reference's SG3Generator wrapper (models/stylegan3/model.py:29-54).
CONFIG_R = dict(channel_base=65536, channel_max=1024, conv_kernel=1, use_radial_filters=True)
CONFIG_T = dict(channel_base=32768, channel_max=512, conv_kernel=3, use_radial_filters=False)


def _shape_is(t, shape):
    assert t.ndim == len(shape) and all(e is None or int(a) == int(e) for a, e in zip(t.shape, shape)), \
        f'wrong tensor shape {tuple(t.shape)}, expected {shape}'


class FullyConnectedLayer(torch.nn.Module):
    """y = act(x @ (weight * weight_gain).T + bias * bias_gain)   (reference :68-100)."""

    def __init__(self, in_features, out_features, activation='linear', bias=True, lr_multiplier=1, weight_init=1,
                 bias_init=0):
        super().__init__()
        self.in_features, self.out_features, self.activation = in_features, out_features, activation
        self.weight = torch.nn.Parameter(torch.randn([out_features, in_features]) * (weight_init / lr_multiplier))
        b0 = np.broadcast_to(np.asarray(bias_init, dtype=np.float32), [out_features]) / lr_multiplier
        self.bias = torch.nn.Parameter(torch.from_numpy(np.ascontiguousarray(b0, dtype=np.float32))) if bias else None
        self.weight_gain = lr_multiplier / np.sqrt(in_features)
        self.bias_gain = lr_multiplier

    def forward(self, x):
        w = self.weight.to(x.dtype) * self.weight_gain
        b = self.bias
        if b is not None:
            b = b.to(x.dtype)
            if self.bias_gain != 1:
                b = b * self.bias_gain
        if self.activation == 'linear' and b is not None:
            return torch.addmm(b.unsqueeze(0), x, w.t())
        return bias_act.bias_act(x.matmul(w.t()), b, act=self.activation)

    def extra_repr(self):
        return f'in_features={self.in_features:d}, out_features={self.out_features:d}, activation={self.activation:s}'


class MappingNetwork(torch.nn.Module):
    """z (and optional label c) -> ws [N, num_ws, w_dim]   (reference :108-160)."""

    def __init__(self, z_dim, c_dim, w_dim, num_ws, num_layers=2, lr_multiplier=0.01, w_avg_beta=0.998):
        super().__init__()
        self.z_dim, self.c_dim, self.w_dim, self.num_ws = z_dim, c_dim, w_dim, num_ws
        self.num_layers, self.w_avg_beta = num_layers, w_avg_beta
        self.embed = FullyConnectedLayer(c_dim, w_dim) if c_dim > 0 else None
        widths = [z_dim + (w_dim if c_dim > 0 else 0)] + [w_dim] * num_layers
        for i in range(num_layers):
            setattr(self, f'fc{i}', FullyConnectedLayer(widths[i], widths[i + 1], activation='lrelu', lr_multiplier=lr_multiplier))
        self.register_buffer('w_avg', torch.zeros([w_dim]))

    @staticmethod
    def _rms_normalize(v):
        return v * (v.square().mean(1, keepdim=True) + 1e-8).rsqrt()

    def forward(self, z, c, truncation_psi=1, truncation_cutoff=None, update_emas=False):
        _shape_is(z, [None, self.z_dim])
        cutoff = self.num_ws if truncation_cutoff is None else truncation_cutoff
        x = self._rms_normalize(z.to(torch.float32))
        if self.c_dim > 0:
            _shape_is(c, [None, self.c_dim])
            x = torch.cat([x, self._rms_normalize(self.embed(c.to(torch.float32)))], dim=1)
        for i in range(self.num_layers):
            x = getattr(self, f'fc{i}')(x)
        if update_emas:
            self.w_avg.copy_(x.detach().mean(dim=0).lerp(self.w_avg, self.w_avg_beta))
        ws = x.unsqueeze(1).repeat([1, self.num_ws, 1])
        if truncation_psi != 1:
            ws[:, :cutoff] = self.w_avg.lerp(ws[:, :cutoff], truncation_psi)
        return ws

    def extra_repr(self):
        return f'z_dim={self.z_dim:d}, c_dim={self.c_dim:d}, w_dim={self.w_dim:d}, num_ws={self.num_ws:d}'


class SynthesisInput(torch.nn.Module):
    """Fourier-feature input with a learned + user-specified transform   (reference :168-249)."""

    def __init__(self, w_dim, channels, size, sampling_rate, bandwidth):
        super().__init__()
        self.w_dim, self.channels = w_dim, channels
        self.size = np.broadcast_to(np.asarray(size), [2])
        self.sampling_rate, self.bandwidth = sampling_rate, bandwidth
        # Frequencies uniform on a disc of radius `bandwidth`, phases uniform in [-0.5, 0.5).
        freqs = torch.randn([channels, 2])
        radii = freqs.square().sum(dim=1, keepdim=True).sqrt()
        freqs /= radii * radii.square().exp().pow(0.25)
        freqs *= bandwidth
        phases = torch.rand([channels]) - 0.5
        self.weight = torch.nn.Parameter(torch.randn([channels, channels]))
        self.affine = FullyConnectedLayer(w_dim, 4, weight_init=0, bias_init=[1, 0, 0, 0])
        self.register_buffer('transform', torch.eye(3, 3))
        self.register_buffer('freqs', freqs)
        self.register_buffer('phases', phases)

    def forward(self, w, t=None):
        if t is None:
            t = self.affine(w)
            t = t / t[:, :2].norm(dim=1, keepdim=True)          # (cos, sin, tx, ty), rotation normalised
        device, n = t.device, t.shape[0]
        rot = torch.eye(3, device=device).repeat(n, 1, 1)
        rot[:, 0, 0], rot[:, 0, 1], rot[:, 1, 0], rot[:, 1, 1] = t[:, 0], -t[:, 1], t[:, 1], t[:, 0]
        trans = torch.eye(3, device=device).repeat(n, 1, 1)
        trans[:, 0, 2], trans[:, 1, 2] = -t[:, 2], -t[:, 3]
        m = rot @ trans @ self.transform                        # rotate, translate, then the user transform
        freqs = self.freqs.unsqueeze(0)
        phases = self.phases.unsqueeze(0) + (freqs @ m[:, :2, 2:]).squeeze(2)
        freqs = freqs @ m[:, :2, :2]
        # Attenuate frequencies pushed beyond the band limit by the user transform.
        amps = (1 - (freqs.norm(dim=2) - self.bandwidth) / (self.sampling_rate / 2 - self.bandwidth)).clamp(0, 1)
        grid = self._sampling_grid(device)
        x = (grid.unsqueeze(3) @ freqs.permute(0, 2, 1).unsqueeze(1).unsqueeze(2)).squeeze(3)     # [n, h, w, c]
        x = torch.sin((x + phases.unsqueeze(1).unsqueeze(2)) * (np.pi * 2)) * amps.unsqueeze(1).unsqueeze(2)
        x = x @ (self.weight / np.sqrt(self.channels)).t()
        x = x.permute(0, 3, 1, 2)
        _shape_is(x, [n, self.channels, int(self.size[1]), int(self.size[0])])
        return x

    def _sampling_grid(self, device):
        """The constant sampling grid of :236-239 (depends on size and sampling rate only).  Built once per device with
        host-side scalars, so the forward itself contains no host-to-device copy and can be captured in a CUDA graph."""
        g = getattr(self, '_grid_cache', None)
        if g is None or g.device != device:
            theta = torch.eye(2, 3)
            theta[0, 0] = 0.5 * self.size[0] / self.sampling_rate
            theta[1, 1] = 0.5 * self.size[1] / self.sampling_rate
            g = torch.nn.functional.affine_grid(theta.unsqueeze(0).to(device), [1, 1, int(self.size[1]), int(self.size[0])],
                                                align_corners=False)
            self._grid_cache = g
        return g

    def extra_repr(self):
        return (f'w_dim={self.w_dim:d}, channels={self.channels:d}, size={list(self.size)},\n'
                f'sampling_rate={self.sampling_rate:g}, bandwidth={self.bandwidth:g}')


class SynthesisLayer(torch.nn.Module):
    """affine -> modulated_conv2d -> filtered_lrelu   (reference :259-401)."""

    def __init__(self, w_dim, is_torgb, is_critically_sampled, use_fp16, in_channels, out_channels, in_size, out_size,
                 in_sampling_rate, out_sampling_rate, in_cutoff, out_cutoff, in_half_width, out_half_width,
                 conv_kernel=3, filter_size=6, lrelu_upsampling=2, use_radial_filters=False, conv_clamp=256,
                 magnitude_ema_beta=0.999):
        super().__init__()
        self.w_dim, self.is_torgb, self.is_critically_sampled, self.use_fp16 = w_dim, is_torgb, is_critically_sampled, use_fp16
        self.in_channels, self.out_channels = in_channels, out_channels
        self.in_size = np.broadcast_to(np.asarray(in_size), [2])
        self.out_size = np.broadcast_to(np.asarray(out_size), [2])
        self.in_sampling_rate, self.out_sampling_rate = in_sampling_rate, out_sampling_rate
        self.tmp_sampling_rate = max(in_sampling_rate, out_sampling_rate) * (1 if is_torgb else lrelu_upsampling)
        self.in_cutoff, self.out_cutoff = in_cutoff, out_cutoff
        self.in_half_width, self.out_half_width = in_half_width, out_half_width
        self.conv_kernel = 1 if is_torgb else conv_kernel
        self.conv_clamp, self.magnitude_ema_beta = conv_clamp, magnitude_ema_beta

        self.affine = FullyConnectedLayer(w_dim, in_channels, bias_init=1)
        self.weight = torch.nn.Parameter(torch.randn([out_channels, in_channels, self.conv_kernel, self.conv_kernel]))
        self.bias = torch.nn.Parameter(torch.zeros([out_channels]))
        self.register_buffer('magnitude_ema', torch.ones([]))

        # The nonlinearity runs at tmp_sampling_rate: upsample in -> tmp, downsample tmp -> out.
        self.up_factor = int(np.rint(self.tmp_sampling_rate / in_sampling_rate))
        assert in_sampling_rate * self.up_factor == self.tmp_sampling_rate
        self.up_taps = filter_size * self.up_factor if self.up_factor > 1 and not is_torgb else 1
        self.register_buffer('up_filter', self.design_lowpass_filter(
            numtaps=self.up_taps, cutoff=in_cutoff, width=in_half_width * 2, fs=self.tmp_sampling_rate))
        self.down_factor = int(np.rint(self.tmp_sampling_rate / out_sampling_rate))
        assert out_sampling_rate * self.down_factor == self.tmp_sampling_rate
        self.down_taps = filter_size * self.down_factor if self.down_factor > 1 and not is_torgb else 1
        self.down_radial = use_radial_filters and not is_critically_sampled
        self.register_buffer('down_filter', self.design_lowpass_filter(
            numtaps=self.down_taps, cutoff=out_cutoff, width=out_half_width * 2, fs=self.tmp_sampling_rate, radial=self.down_radial))

        # Padding so that the output has exactly out_size samples centred on the input grid.
        total = (self.out_size - 1) * self.down_factor + 1
        total = total - (self.in_size + self.conv_kernel - 1) * self.up_factor + self.up_taps + self.down_taps - 2
        lo = (total + self.up_factor) // 2
        hi = total - lo
        self.padding = [int(lo[0]), int(hi[0]), int(lo[1]), int(hi[1])]

    def forward(self, x, w, styles=None, noise_mode='random', force_fp32=False, update_emas=False):
        x = self.conv_part(x, w, styles=styles, noise_mode=noise_mode, force_fp32=force_fp32, update_emas=update_emas)
        return self.act_part(x)

    # The two halves of forward(), separately callable so that `PipelinedSynthesis` can launch them on different streams.
    def conv_part(self, x, w, styles=None, noise_mode='random', force_fp32=False, update_emas=False):
        """EMA gain -> affine -> modulated_conv2d   (reference :335-358)."""
        assert noise_mode in ['random', 'const', 'none']
        _shape_is(x, [None, self.in_channels, int(self.in_size[1]), int(self.in_size[0])])
        if update_emas:
            with torch.autograd.profiler.record_function('update_magnitude_ema'):
                cur = x.detach().to(torch.float32).square().mean()
                self.magnitude_ema.copy_(cur.lerp(self.magnitude_ema, self.magnitude_ema_beta))
        input_gain = self.magnitude_ema.rsqrt()
        if styles is None:
            _shape_is(w, [x.shape[0], self.w_dim])
            styles = self.affine(w)
            if self.is_torgb:
                styles = styles * (1 / np.sqrt(self.in_channels * (self.conv_kernel ** 2)))
        dtype = torch.float16 if (self.use_fp16 and not force_fp32 and x.device.type == 'cuda') else torch.float32
        return modulated_conv2d(x=x.to(dtype), w=self.weight, s=styles, padding=self.conv_kernel - 1,
                                demodulate=(not self.is_torgb), input_gain=input_gain)

    def act_part(self, x):
        """bias -> filtered leaky ReLU at the temporary sampling rate   (reference :361-368)."""
        dtype = x.dtype
        # policy 'round': the next layer's convolution reads this output with TF32 tensor cores (which truncate): round to nearest
        # here.  (The default policy compensates the truncation in the conv's weight prologue instead, see modulated_conv.py.)
        rnd = (not self.is_torgb) and dtype == torch.float32 and _math_mode() == 'tf32' and tf32_activation_policy() == 'round'
        with filtered_lrelu.tf32_rounded_outputs(rnd):
            x = filtered_lrelu.filtered_lrelu(
                x=x, fu=self.up_filter, fd=self.down_filter, b=self.bias.to(x.dtype), up=self.up_factor, down=self.down_factor,
                padding=self.padding, gain=(1 if self.is_torgb else np.sqrt(2)), slope=(1 if self.is_torgb else 0.2),
                clamp=self.conv_clamp)
        _shape_is(x, [None, self.out_channels, int(self.out_size[1]), int(self.out_size[0])])
        assert x.dtype == dtype
        return x

    @staticmethod
    def design_lowpass_filter(numtaps, cutoff, width, fs, radial=False):
        """Kaiser-windowed low-pass taps: separable (firwin) or radially symmetric (jinc), float32."""
        assert numtaps >= 1
        if numtaps == 1:
            return None
        if not radial:
            return torch.as_tensor(scipy.signal.firwin(numtaps=numtaps, cutoff=cutoff, width=width, fs=fs), dtype=torch.float32)
        pos = (np.arange(numtaps) - (numtaps - 1) / 2) / fs
        rad = np.hypot(*np.meshgrid(pos, pos))
        taps = scipy.special.j1(2 * cutoff * (np.pi * rad)) / (np.pi * rad)
        win = np.kaiser(numtaps, scipy.signal.kaiser_beta(scipy.signal.kaiser_atten(numtaps, width / (fs / 2))))
        taps *= np.outer(win, win)
        taps /= np.sum(taps)
        return torch.as_tensor(taps, dtype=torch.float32)

    def extra_repr(self):
        return (f'w_dim={self.w_dim:d}, is_torgb={self.is_torgb}, is_critically_sampled={self.is_critically_sampled}, '
                f'use_fp16={self.use_fp16},\nin_size={list(self.in_size)}, out_size={list(self.out_size)}, '
                f'in_channels={self.in_channels:d}, out_channels={self.out_channels:d}, up={self.up_factor}, down={self.down_factor}')


class SynthesisNetwork(torch.nn.Module):
    """Fourier input + num_layers synthesis layers + ToRGB   (reference :405-525)."""

    def __init__(self, w_dim, img_resolution, img_channels, channel_base=32768, channel_max=512, num_layers=14,
                 num_critical=2, first_cutoff=2, first_stopband=2 ** 2.1, last_stopband_rel=2 ** 0.3, margin_size=10,
                 output_scale=0.25, num_fp16_res=4, **layer_kwargs):
        super().__init__()
        self.w_dim, self.num_ws = w_dim, num_layers + 2
        self.img_resolution, self.img_channels = img_resolution, img_channels
        self.num_layers, self.num_critical = num_layers, num_critical
        self.margin_size, self.output_scale, self.num_fp16_res = margin_size, output_scale, num_fp16_res

        # Cutoffs and stopbands grow geometrically over the non-critical layers, then stay flat.
        last_cutoff = img_resolution / 2
        last_stopband = last_cutoff * last_stopband_rel
        expo = np.minimum(np.arange(num_layers + 1) / (num_layers - num_critical), 1)
        cutoffs = first_cutoff * (last_cutoff / first_cutoff) ** expo
        stopbands = first_stopband * (last_stopband / first_stopband) ** expo
        rates = np.exp2(np.ceil(np.log2(np.minimum(stopbands * 2, img_resolution))))
        half_widths = np.maximum(stopbands, rates / 2) - cutoffs
        sizes = rates + margin_size * 2
        sizes[-2:] = img_resolution
        channels = np.rint(np.minimum((channel_base / 2) / cutoffs, channel_max))
        channels[-1] = img_channels

        self.input = SynthesisInput(w_dim=w_dim, channels=int(channels[0]), size=int(sizes[0]),
                                    sampling_rate=rates[0], bandwidth=cutoffs[0])
        self.layer_names = []
        for i in range(num_layers + 1):
            p = max(i - 1, 0)
            layer = SynthesisLayer(
                w_dim=w_dim, is_torgb=(i == num_layers), is_critically_sampled=(i >= num_layers - num_critical),
                use_fp16=bool(rates[i] * (2 ** num_fp16_res) > img_resolution),
                in_channels=int(channels[p]), out_channels=int(channels[i]), in_size=int(sizes[p]), out_size=int(sizes[i]),
                in_sampling_rate=int(rates[p]), out_sampling_rate=int(rates[i]), in_cutoff=cutoffs[p], out_cutoff=cutoffs[i],
                in_half_width=half_widths[p], out_half_width=half_widths[i], **layer_kwargs)
            name = f'L{i}_{layer.out_size[0]}_{layer.out_channels}'
            setattr(self, name, layer)
            self.layer_names.append(name)

    def forward(self, ws, all_s=None, **layer_kwargs):
        if all_s is None:
            _shape_is(ws, [None, self.num_ws, self.w_dim])
            ws = ws.to(torch.float32).unbind(dim=1)
            x = self.input(ws[0])
            for name, w in zip(self.layer_names, ws[1:]):
                x = getattr(self, name)(x, w, **layer_kwargs)
        else:                                                   # S-space path (fork addition, reference :481-486)
            x = self.input(None, t=all_s['input'])
            for name in self.layer_names:
                x = getattr(self, name)(x, None, styles=all_s[name], **layer_kwargs)
        if self.output_scale != 1:
            x = x * self.output_scale
        _shape_is(x, [None, self.img_channels, self.img_resolution, self.img_resolution])
        return x.to(torch.float32)

    def W2S(self, ws):
        """W+ codes -> per-layer style vectors (S space), keyed like `all_s` of forward()."""
        _shape_is(ws, [None, self.num_ws, self.w_dim])
        ws = ws.to(torch.float32).unbind(dim=1)
        t = self.input.affine(ws[0])
        all_s = {'input': t / t[:, :2].norm(dim=1, keepdim=True)}
        for name, w in zip(self.layer_names, ws[1:]):
            layer = getattr(self, name)
            styles = layer.affine(w)
            if layer.is_torgb:
                styles = styles * (1 / np.sqrt(layer.in_channels * (layer.conv_kernel ** 2)))
            all_s[name] = styles
        return all_s

    def extra_repr(self):
        return (f'w_dim={self.w_dim:d}, num_ws={self.num_ws:d}, img_resolution={self.img_resolution:d}, '
                f'img_channels={self.img_channels:d}, num_layers={self.num_layers:d}, num_critical={self.num_critical:d}')


class Generator(torch.nn.Module):
    def __init__(self, z_dim, c_dim, w_dim, img_resolution, img_channels, mapping_kwargs={}, **synthesis_kwargs):
        super().__init__()
        self.z_dim, self.c_dim, self.w_dim = z_dim, c_dim, w_dim
        self.img_resolution, self.img_channels = img_resolution, img_channels
        self.synthesis = SynthesisNetwork(w_dim=w_dim, img_resolution=img_resolution, img_channels=img_channels, **synthesis_kwargs)
        self.num_ws = self.synthesis.num_ws
        self.mapping = MappingNetwork(z_dim=z_dim, c_dim=c_dim, w_dim=w_dim, num_ws=self.num_ws, **mapping_kwargs)

    def forward(self, z, c, truncation_psi=1, truncation_cutoff=None, update_emas=False, **synthesis_kwargs):
        ws = self.mapping(z, c, truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff, update_emas=update_emas)
        return self.synthesis(ws, update_emas=update_emas, **synthesis_kwargs)


class GraphedSynthesis:
    """CUDA-graph replay of `SynthesisNetwork.forward` for one (batch, dtype) signature.

    A forward of the 1024^2 generator is ~300 kernel launches of which only ~45 are heavy; at batch 1-8 (ReStyle inversion,
    PTI, video frames) the Python / launch overhead of the ~250 small ones is comparable to the GPU time (SURVEY.md 8f,
    rank 2).  This wrapper captures one eager forward into a `torch.cuda.CUDAGraph` -- every sg3_b200 kernel launches on
    the capturing stream, taps and TMA tensor maps travel in the launch parameters, nothing synchronises with the host --
    and replays it: `img = graphed(ws)` copies `ws` into the static input and returns the static output (valid until the
    next call).  Inference only (`torch.no_grad`); the generator's weights are read at replay time, so in-place weight
    updates (PTI) are seen, but `input.transform` must be updated in place as well.
    """

    def __init__(self, synthesis, ws_example, **layer_kwargs):
        self.synthesis = synthesis
        self.kw = dict(noise_mode='const', force_fp32=True)
        self.kw.update(layer_kwargs)
        self.ws = ws_example.detach().clone()
        side = torch.cuda.Stream(self.ws.device)
        side.wait_stream(torch.cuda.current_stream(self.ws.device))
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(2):                                  # warm-up: filter tap caches, allocator pools, lazy inits
                self.synthesis(self.ws, **self.kw)
        torch.cuda.current_stream(self.ws.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.img = self.synthesis(self.ws, **self.kw)

    def __call__(self, ws):
        assert ws.shape == self.ws.shape and ws.dtype == self.ws.dtype
        self.ws.copy_(ws, non_blocking=True)
        self.graph.replay()
        return self.img


class PipelinedSynthesis:
    """`SynthesisNetwork.forward` with the two halves of every layer OVERLAPPED on the GPU.

    A synthesis layer is a modulated convolution (tcgen05 tensor cores, bound by HBM in config R and by the tensor pipe in config
    T; its FP32 pipe idles) followed by the fused filtered_lrelu stencil (bound by the FP32 pipe; HBM at ~20 %, tensor pipe idle).
    Run back to back, each leaves the other's resource unused.  Here the batch is split into micro-batches and the network is
    software-pipelined over two streams: the convolutions go to a HIGH-priority stream, the stencils to a normal one, chained per
    micro-batch by events -- while the stencil of layer k works on micro-batch m, the convolution of micro-batch m+1 (or of layer
    k+1 on micro-batch m-1) may run.  With `conv_smem_budget=CONV_SMEM_BUDGET` co-residency is arranged, not hoped for: the conv
    kernels are persistent (one CTA per SM) and get a shared-memory budget (`sg3_modconv_set_smem_budget`) that lets one conv CTA
    (192 threads, 64 registers) sit next to three of the four stencil CTAs an SM holds (4 warps, 128 registers, 28 KB each), and the
    block scheduler serves the high-priority stream first, so a pending conv CTA takes the first stencil slot that retires.  The
    residency trace confirms that this happens -- and that it buys nothing (profiles/r02_overlap.md), so the default keeps the full
    conv kernels and only overlaps kernel tails and launch gaps (+2 % on config R at batch 32).

    Every op is per sample, so the result equals the plain forward up to the batch-global style RMS of networks_stylegan3.py:42,
    which cancels under demodulation except for its 1e-8 epsilon (same statement as for batch sharding across GPUs).
    Inference only (`torch.no_grad`)."""

    CONV_SMEM_BUDGET = 138 * 1024       # 228 KB per SM - 3 x (28 KB + 1 KB reserved) stencil CTAs - 1 KB reserved - barriers

    def __init__(self, synthesis, micro_batches=2, conv_smem_budget=0):
        # conv_smem_budget: 0 (default) = the conv kernels keep their full TMA ring, the two streams overlap at kernel tails only;
        # CONV_SMEM_BUDGET = one conv CTA fits next to three stencil CTAs per SM (true co-residency).  Measured on B200
        # (profiles/r02_overlap.md): co-residency is never faster -- both kernels lean on the shared-memory data pipe -- so it is off.
        self.synthesis = synthesis
        self.micro_batches = int(micro_batches)
        self.budget = int(conv_smem_budget)
        self._streams = {}

    def _get_streams(self, device):
        key = torch.device(device).index
        if key not in self._streams:
            try:
                lo_pri, hi_pri = torch.cuda.Stream.priority_range()        # (least, greatest) = (0, -5) on B200
            except Exception:
                lo_pri, hi_pri = 0, -1
            self._streams[key] = (torch.cuda.Stream(device, priority=hi_pri), torch.cuda.Stream(device, priority=lo_pri))
        return self._streams[key]

    @torch.no_grad()
    def __call__(self, ws, out=None, **layer_kwargs):
        from . import capi
        syn = self.synthesis
        _shape_is(ws, [None, syn.num_ws, syn.w_dim])
        N, dev = ws.shape[0], ws.device
        M = max(1, min(self.micro_batches, N))
        if M == 1:
            img = syn(ws, **layer_kwargs)
            if out is not None:
                out.copy_(img)
                return out
            return img
        if out is None:
            out = torch.empty([N, syn.img_channels, syn.img_resolution, syn.img_resolution], dtype=torch.float32, device=dev)
        bounds = [(N * m // M, N * (m + 1) // M) for m in range(M)]
        conv_s, act_s = self._get_streams(dev)
        main = torch.cuda.current_stream(dev)
        conv_s.wait_stream(main)
        act_s.wait_stream(main)
        layers = [getattr(syn, name) for name in syn.layer_names]
        prev = capi.lib().sg3_modconv_set_smem_budget(self.budget)
        try:
            # Tensor lifetimes across the two streams are ordered by hand instead of `Tensor.record_stream` (which parks a freed
            # block until the GPU has caught up -- the host runs a whole forward ahead, so every activation of the network would
            # be live at once): a conv output t is allocated on the conv stream and read by the stencil stream; it is released only
            # after the conv stream has waited for that stencil (the wait the next layer needs anyway), so any later conv-stream
            # allocation that reuses the block is ordered behind the read.  Symmetrically for the stencil outputs.
            wsm, x, t_prev, ev = [], [None] * M, [None] * M, [None] * M
            with torch.cuda.stream(conv_s):
                for m, (b0, b1) in enumerate(bounds):
                    w = ws[b0:b1].to(torch.float32).unbind(dim=1)
                    wsm.append(w)
                    x[m] = syn.input(w[0])
            for k, layer in enumerate(layers):
                for m in range(M):
                    with torch.cuda.stream(conv_s):
                        if ev[m] is not None:
                            conv_s.wait_event(ev[m])                     # the stencil of the previous layer on this micro-batch ...
                            t_prev[m] = None                             # ... which was the last reader of that layer's conv output
                        t = layer.conv_part(x[m], wsm[m][k + 1], **layer_kwargs)
                        done = torch.cuda.Event()
                        done.record(conv_s)
                    with torch.cuda.stream(act_s):
                        act_s.wait_event(done)
                        x[m] = None                                      # the conv that read it is complete as far as act_s is concerned
                        y = layer.act_part(t)
                        if k == len(layers) - 1:                         # output scale and fp32 cast of SynthesisNetwork.forward
                            b0, b1 = bounds[m]
                            torch.mul(y, syn.output_scale, out=out[b0:b1])
                        ev[m] = torch.cuda.Event()
                        ev[m].record(act_s)
                    x[m], t_prev[m] = y, t
                    del t, y
            conv_s.wait_stream(act_s)                                    # orders the release of the last conv outputs (below)
            act_s.wait_stream(conv_s)
            del x, t_prev
        finally:
            capi.lib().sg3_modconv_set_smem_budget(prev)
            main.wait_stream(act_s)                                      # also on an exception: later work on `main` stays ordered
            main.wait_stream(conv_s)
        return out
