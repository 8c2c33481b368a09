"""Harness models for BASELINE.json configs[3] and configs[4]: the CALLERS of the synthesis hot path, restated at the shape the
reference uses so that `bench.py --config restyle | pti` can measure the hot path inside its real loops.

Nothing here is product code of the hot path: the encoder and the perceptual network are plain PyTorch modules (cuDNN,
channels_last, the reference's fp32 / TF32 arithmetic), random-init because no checkpoint can be fetched; the generator calls
are `sg3_b200.networks` (the kernels under test).  What is mirrored, with the reference lines:

  * ReStyle-pSp encoder: `BackboneEncoder(50, 'ir_se', n_styles=16, input_nc=6)`
    (models/setgan/encoder/encoders/restyle_psp_encoders.py:9-51, helpers.py:21-140, map2style.py:8-27): IR-SE50 body
    [3, 4, 14, 3] bottleneck_IR_SE units on a 256^2 6-channel input, 16 GradualStyleBlock heads on the final 16x16 map.
  * pSp forward + iterative refinement: models/setgan/encoder/psp3.py:44-83 (codes = encoder(x) + latent | latent_avg;
    synthesis with the identity transform; a second synthesis with the per-sample [N, 3, 3] landmarks transforms; `face_pool`
    1024 -> 256) and utils/inference_utils.py:67-111 (`run_on_batch`: n_iters_per_batch = 5, inputs concatenated with the
    previous output, last iteration returns the unaligned image).
  * PTI: inversion/video/run_pti_video.py:96-168 + inversion/scripts/run_pti_images.py:167-177: Adam(lr 3e-4) on
    list(G.synthesis.parameters())[3:], loss = l2_lambda * MSE + lpips_lambda * LPIPS(alex) (criteria/lpips/lpips.py:8-35,
    networks.py:21-83), batch 4 frames per GPU; data-parallel over frames with one flat-bucket NCCL all-reduce per step
    (sg3_b200.sharding.FlatGradBucket, the pattern of setgan/training_loop.py:446-455).
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------
# IR-SE50 ReStyle encoder (random init)

class _SE(nn.Module):
    def __init__(self, c, r=16):
        super().__init__()
        self.fc1 = nn.Conv2d(c, c // r, 1, bias=False)
        self.fc2 = nn.Conv2d(c // r, c, 1, bias=False)

    def forward(self, x):
        s = F.adaptive_avg_pool2d(x, 1)
        return x * torch.sigmoid(self.fc2(F.relu(self.fc1(s))))


class _BottleneckIRSE(nn.Module):
    def __init__(self, cin, depth, stride):
        super().__init__()
        if cin == depth:
            self.shortcut = nn.MaxPool2d(1, stride)
        else:
            self.shortcut = nn.Sequential(nn.Conv2d(cin, depth, 1, stride, bias=False), nn.BatchNorm2d(depth))
        self.res = nn.Sequential(
            nn.BatchNorm2d(cin), nn.Conv2d(cin, depth, 3, 1, 1, bias=False), nn.PReLU(depth),
            nn.Conv2d(depth, depth, 3, stride, 1, bias=False), nn.BatchNorm2d(depth), _SE(depth, 16))

    def forward(self, x):
        return self.res(x) + self.shortcut(x)


class _StyleHead(nn.Module):
    """GradualStyleBlock(512, 512, 16): four stride-2 3x3 convs + LeakyReLU, then an equalised linear layer."""

    def __init__(self, cin=512, cout=512, spatial=16):
        super().__init__()
        mods, c = [], cin
        for _ in range(int(math.log2(spatial))):
            mods += [nn.Conv2d(c, cout, 3, 2, 1), nn.LeakyReLU()]
            c = cout
        self.convs = nn.Sequential(*mods)
        self.weight = nn.Parameter(torch.randn(cout, cout))
        self.bias = nn.Parameter(torch.zeros(cout))
        self.scale = 1 / math.sqrt(cout)

    def forward(self, x):
        x = self.convs(x).flatten(1)
        return F.linear(x, self.weight * self.scale, self.bias)


class RestyleEncoder(nn.Module):
    def __init__(self, n_styles=16, input_nc=6):
        super().__init__()
        self.input_layer = nn.Sequential(nn.Conv2d(input_nc, 64, 3, 1, 1, bias=False), nn.BatchNorm2d(64), nn.PReLU(64))
        units, cin = [], 64
        for depth, n in ((64, 3), (128, 4), (256, 14), (512, 3)):
            for i in range(n):
                units.append(_BottleneckIRSE(cin, depth, 2 if i == 0 else 1))
                cin = depth
        self.body = nn.Sequential(*units)
        self.styles = nn.ModuleList([_StyleHead() for _ in range(n_styles)])

    def forward(self, x):
        x = self.body(self.input_layer(x))
        return torch.stack([s(x) for s in self.styles], dim=1)          # [N, n_styles, 512]


# ------------------------------------------------------------------------------------------------
# pSp wrapper and the iterative inversion loop

class PSP:
    """Encoder + StyleGAN3 decoder, inference only (psp3.py:44-83).  `use_graph=True` replays the two synthesis calls of an
    iteration as CUDA graphs (sg3_b200.networks.GraphedSynthesis; the per-sample transforms are updated in place)."""

    def __init__(self, encoder, G, use_graph=True):
        self.encoder, self.G = encoder, G
        self.latent_avg = G.mapping.w_avg.detach()
        self.use_graph = use_graph
        self._graphs = {}

    def face_pool(self, img):
        return F.adaptive_avg_pool2d(img, (256, 256))

    def _synthesis(self, codes, transform):
        syn = self.G.synthesis
        n = codes.shape[0]
        if tuple(syn.input.transform.shape) != (n, 3, 3):
            syn.input.transform = torch.eye(3, device=codes.device).repeat(n, 1, 1)
            self._graphs.clear()
        syn.input.transform.copy_(transform)
        if not self.use_graph:
            return syn(codes, noise_mode='const', force_fp32=True)
        key = (n, codes.dtype)
        if key not in self._graphs:
            from sg3_b200 import networks
            self._graphs[key] = networks.GraphedSynthesis(syn, codes)
            syn.input.transform.copy_(transform)
        return self._graphs[key](codes)

    @torch.no_grad()
    def forward(self, x, latent=None, landmarks_transform=None):
        codes = self.encoder(x.contiguous(memory_format=torch.channels_last)).float()
        codes = codes + (latent if latent is not None else self.latent_avg.expand(codes.shape[0], codes.shape[1], -1))
        n = x.shape[0]
        ident = torch.eye(3, device=x.device).expand(n, 3, 3)
        images = self.face_pool(self._synthesis(codes, ident))
        unaligned = None
        if landmarks_transform is not None:
            unaligned = self.face_pool(self._synthesis(codes, landmarks_transform.float()))
        return images, unaligned, codes


@torch.no_grad()
def run_on_batch(inputs, net, avg_image, n_iters=5, landmarks_transform=None):
    """utils/inference_utils.py:67-111; returns the final image [N, 3, 256, 256] and latent [N, 16, 512]."""
    y_hat = latent = None
    for it in range(n_iters):
        prev = avg_image.unsqueeze(0).expand(inputs.shape[0], -1, -1, -1) if it == 0 else y_hat
        x_input = torch.cat([inputs, prev], dim=1)
        images, unaligned, latent = net.forward(x_input, latent=latent, landmarks_transform=landmarks_transform)
        last = it == n_iters - 1
        y_hat = unaligned if (landmarks_transform is not None and last) else images
    return y_hat, latent


def random_landmarks_transforms(n, generator, device):
    """Per-sample similarity transforms of the size the video alignment produces: small rotation, scale, translation."""
    ang = (torch.rand(n, generator=generator) - 0.5) * 0.3
    sc = 1 + (torch.rand(n, generator=generator) - 0.5) * 0.1
    t = (torch.rand(n, 2, generator=generator) - 0.5) * 0.1
    m = torch.eye(3).repeat(n, 1, 1)
    m[:, 0, 0], m[:, 0, 1] = sc * torch.cos(ang), -sc * torch.sin(ang)
    m[:, 1, 0], m[:, 1, 1] = sc * torch.sin(ang), sc * torch.cos(ang)
    m[:, :2, 2] = t
    return m.to(device)


# ------------------------------------------------------------------------------------------------
# LPIPS (AlexNet structure, random init) and the PTI step

class LPIPSAlex(nn.Module):
    """criteria/lpips: z-score, AlexNet features at the five ReLUs, channel-normalised, squared difference, 1x1 'lin' layers,
    spatial mean, sum.  Random-init torchvision AlexNet and random non-negative lin weights (no pretrained weights offline)."""

    def __init__(self):
        super().__init__()
        import torchvision
        self.layers = torchvision.models.alexnet(weights=None).features
        self.targets = (2, 5, 8, 10, 12)
        self.lin = nn.ModuleList([nn.Conv2d(c, 1, 1, bias=False) for c in (64, 192, 384, 256, 256)])
        for l in self.lin:
            l.weight.data.uniform_(0, 1e-2)
        self.register_buffer('mean', torch.tensor([-.030, -.088, -.188])[None, :, None, None])
        self.register_buffer('std', torch.tensor([.458, .448, .450])[None, :, None, None])
        self.requires_grad_(False)

    def features(self, x):
        x = (x - self.mean) / self.std
        out = []
        for i, layer in enumerate(self.layers, 1):
            x = layer(x)
            if i in self.targets:
                out.append(x / (x.square().sum(1, keepdim=True).sqrt() + 1e-10))
            if len(out) == len(self.targets):
                break
        return out

    def forward(self, x, y):
        res = [l((fx - fy) ** 2).mean((2, 3), True) for fx, fy, l in zip(self.features(x), self.features(y), self.lin)]
        return torch.sum(torch.cat(res, 0)) / x.shape[0]


class PTITrainer:
    """Data-parallel PTI over frames.  Every rank owns a contiguous shard of the frames and walks it in batches of
    `batch` frames; one `step()` = forward, loss, backward, ONE flat all-reduce, Adam.

    `use_graph=True`: a PTI step at batch 4 is ~800 kernel launches (autograd through 15 layers, ~60 parameter tensors), and the
    Python / launch overhead of the small ones leaves the GPU idle for a third of the step.  The step is therefore captured ONCE
    into two CUDA graphs -- (zero the gradient bucket, forward, loss, backward) and (Adam) -- around the eager NCCL all-reduce, and
    replayed on static input buffers.  Every sg3_b200 kernel launches on the capturing stream, taps and tensor maps travel in the
    launch parameters, nothing synchronises with the host."""

    def __init__(self, G, frames, latents, batch=4, lr=3e-4, l2_lambda=1.0, lpips_lambda=1.0, world=1, use_graph=False):
        from sg3_b200 import sharding
        self.G, self.frames, self.latents, self.batch, self.world = G, frames, latents, batch, world
        self.l2_lambda, self.lpips_lambda = l2_lambda, lpips_lambda
        self.lpips = LPIPSAlex().to(frames.device).eval()
        params = list(G.synthesis.parameters())[3:]            # as run_pti_images.py: skip the Fourier-feature input
        for p in G.parameters():
            p.requires_grad_(False)
        for p in params:
            p.requires_grad_(True)
        self.bucket = sharding.FlatGradBucket(params)
        self.opt = torch.optim.Adam(params, lr=lr, capturable=use_graph)
        self.cursor = 0
        self.use_graph = use_graph
        self._graphs = None
        dev = frames.device
        self.s_tgt = torch.empty([batch] + list(frames.shape[1:]), device=dev)       # static inputs of the captured step
        self.s_ws = torch.empty([batch] + list(latents.shape[1:]), device=dev)
        self.s_loss = torch.zeros([], device=dev)

    def _next_batch(self):
        n = self.frames.shape[0]
        idx = [(self.cursor + i) % n for i in range(self.batch)]
        self.cursor = (self.cursor + self.batch) % n
        return idx

    def _fwd_bwd(self):
        self.bucket.zero()
        img = self.G.synthesis(self.s_ws, noise_mode='const', force_fp32=True)
        loss = self.l2_lambda * F.mse_loss(img, self.s_tgt) + self.lpips_lambda * self.lpips(img, self.s_tgt)
        loss.backward()
        self.s_loss.copy_(loss.detach())

    def _capture(self):
        side = torch.cuda.Stream(self.s_tgt.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                          # warm-up off the default stream: allocator pools, tap caches, Adam state
            for _ in range(3):
                self._fwd_bwd()
                self.bucket.all_reduce_mean()
                self.opt.step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1):
            self._fwd_bwd()
        with torch.cuda.graph(g2):
            self.opt.step()
        self._graphs = (g1, g2)

    def load_batch(self, tgt, ws):
        self.s_tgt.copy_(tgt, non_blocking=True)
        self.s_ws.copy_(ws, non_blocking=True)

    def run_loaded(self):
        """One optimisation step on the batch currently in the static buffers; returns the (device) loss."""
        if self.use_graph:
            if self._graphs is None:
                self._capture()
            self._graphs[0].replay()
            self.bucket.all_reduce_mean()
            self._graphs[1].replay()
        else:
            self._fwd_bwd()
            self.bucket.all_reduce_mean()
            self.opt.step()
        return self.s_loss

    def step(self):
        idx = self._next_batch()
        self.load_batch(self.frames[idx], self.latents[idx])
        return self.run_loaded().clone()
