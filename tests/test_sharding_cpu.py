"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: batch sharding with no collective on the data
path, max-over-ranks timing, parameter broadcast and the flat-bucket gradient all-reduce of the DP step."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import sg3_b200  # noqa: F401
    from sg3_b200 import sharding
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        assert sharding.rank_info() == (rank, world, rank)
        # 1. batch sharding: contiguous, disjoint, complete
        batch = torch.arange(7 * 3, dtype=torch.float32).reshape(7, 3)
        mine = sharding.shard_batch(batch, rank, world)
        gathered = [torch.zeros(4, 3) for _ in range(world)]
        pad = torch.zeros(4, 3)
        pad[:mine.shape[0]] = mine
        dist.all_gather(gathered, pad)
        sizes = [sharding.shard_range(7, r, world) for r in range(world)]
        rebuilt = torch.cat([g[:e - b] for g, (b, e) in zip(gathered, sizes)])
        assert torch.equal(rebuilt, batch)
        # 2. timing reduction: the slowest rank wins
        assert sharding.max_over_ranks(10.0 + rank) == 10.0 + world - 1
        # 3. DP step: broadcast params, per-rank grads, one all-reduce over a flat bucket
        torch.manual_seed(100 + rank)                     # ranks start different ...
        model = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.Linear(4, 2))
        sharding.broadcast_parameters(model)              # ... and end equal
        ref = [p.detach().clone() for p in model.parameters()]
        bucket = sharding.FlatGradBucket(model.parameters())
        assert bucket.numel == sum(p.numel() for p in model.parameters())
        torch.manual_seed(7)
        frames = torch.randn(6, 5)
        target = torch.randn(6, 2)
        x, t = sharding.shard_batch(frames, rank, world), sharding.shard_batch(target, rank, world)
        bucket.zero()
        loss = ((model(x) - t) ** 2).sum() / frames.shape[0] * world     # mean over the global batch after /world
        loss.backward()
        assert model[0].weight.grad.data_ptr() == bucket.flat.data_ptr()          # grads live in the bucket
        bucket.all_reduce_mean()
        # single-process reference on the full batch
        single = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.Linear(4, 2))
        for p, r in zip(single.parameters(), ref):
            p.data.copy_(r)
        ((single(frames) - target) ** 2).sum().div(frames.shape[0]).backward()
        for p, q in zip(model.parameters(), single.parameters()):
            assert torch.allclose(p.grad, q.grad, atol=1e-6), (p.grad - q.grad).abs().max()
        # NaN guard
        bucket.flat[0] = float('nan')
        bucket.all_reduce_mean()
        assert torch.isfinite(bucket.flat).all()
        out.put((rank, 'ok'))
    except Exception as e:  # pragma: no cover
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    import sys
    from conftest import ROOT
    sys.path.insert(0, ROOT)
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, 'ok'), (1, 'ok')], results


def test_shard_range_properties():
    import sg3_b200  # noqa: F401
    from sg3_b200 import sharding
    for n in (0, 1, 5, 8, 64, 65):
        for w in (1, 2, 3, 8):
            r = [sharding.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
            assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1
