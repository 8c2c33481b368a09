"""Multi-GPU plumbing for the synthesis path: one process per GPU, batch sharding with no data-path
collective for forward / inversion, and a flat-bucket gradient all-reduce for data-parallel PTI /
encoder fine-tuning (the pattern of the reference's only distributed code, setgan/training_loop.py:446-455,
applied to the generator's parameters).  Works with any torch.distributed backend (NCCL on GPUs, gloo in
the CPU tests)."""
import os

import torch
import torch.distributed as dist


def rank_info():
    """(rank, world_size, local_rank) from the torchrun environment (single process -> (0, 1, 0))."""
    return int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('LOCAL_RANK', '0'))


def shard_range(num_items, rank, world_size):
    """Contiguous [begin, end) of `num_items` samples owned by `rank`; sizes differ by at most one."""
    assert 0 <= rank < world_size
    base, extra = divmod(num_items, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_batch(tensor, rank, world_size):
    """The rank's contiguous slice of a batch-major tensor (latents, frames, targets)."""
    b, e = shard_range(tensor.shape[0], rank, world_size)
    return tensor[b:e]


def max_over_ranks(value, device='cpu'):
    """Max of a python float over all ranks (timing: the slowest rank defines the step)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def broadcast_parameters(module, src=0):
    """Make every rank start from rank `src`'s parameters and buffers (setgan/training_loop.py:278-281)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src)


class FlatGradBucket:
    """One contiguous fp32 buffer holding the gradients of `params`; `.grad` of every parameter is a view
    into it, so backward writes straight into the bucket and the data-parallel step is a single all-reduce
    (no torch.cat copy as in setgan/training_loop.py:449)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        assert self.params, 'no trainable parameters'
        dev = self.params[0].device
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        ofs = 0
        for p in self.params:
            assert p.dtype == torch.float32 and p.device == dev
            p.grad = self.flat[ofs:ofs + p.numel()].view_as(p)
            ofs += p.numel()

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self):
        """sum over ranks / world size, then the reference's NaN/Inf guard (training_loop.py:452)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(dist.get_world_size())
        torch.nan_to_num(self.flat, nan=0.0, posinf=1e5, neginf=-1e5, out=self.flat)
        return self.flat


class HostPipeline:
    """Generator forward with HOST inputs and outputs, copies overlapped with compute.

    `submit(ws_host, img_host)` enqueues: pinned latents -> device (compute stream), `G.synthesis`, then the device -> host
    copy of the images on a side stream, so the copy of batch i runs under the forward of batch i+1 (the images of a
    1024^2 batch are 12.6 MB each; a serial copy costs 5-10 % of a step).  `img_host` may be reused once `finish()` or the
    next-but-one `submit()` returned.  Inference / inversion loops that keep their frames on the host use this; nothing
    here changes what is computed."""

    def __init__(self, G, device, **synthesis_kwargs):
        self.G = G
        self.device = torch.device(device)
        self.kw = dict(noise_mode='const', force_fp32=True)
        self.kw.update(synthesis_kwargs)
        self.copy_stream = torch.cuda.Stream(self.device)
        self._pending = []          # (event, device tensor) of copies in flight

    def submit(self, ws_host, img_host):
        cur = torch.cuda.current_stream(self.device)
        with torch.no_grad():
            ws = ws_host.to(self.device, non_blocking=True)
            img = self.G.synthesis(ws, **self.kw)
        ready = torch.cuda.Event()
        ready.record(cur)
        while len(self._pending) >= 2:                      # bound the device memory held by copies in flight
            ev, _ = self._pending.pop(0)
            ev.synchronize()
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(ready)
            img_host.copy_(img, non_blocking=True)
            img.record_stream(self.copy_stream)
            done = torch.cuda.Event()
            done.record(self.copy_stream)
        self._pending.append((done, img))
        return img_host

    def finish(self):
        self.copy_stream.synchronize()
        self._pending.clear()
