"""Multi-rank host logic on CPU: world_size-2 gloo run of imageclassification_b200.ddp.DistributedDataParallel.
Checks: rank-0 parameters are broadcast at construction; after backward every rank holds the MEAN of
the per-rank gradients, equal to single-process training on the concatenated batch within fp32 reassociation;
gradients live in the flat arena; `zero_grad()` (set_to_none) and gradient accumulation (no no_sync, as the reference)
both keep working; the reducer's bucket layout covers every parameter exactly once."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model(seed):
    torch.manual_seed(seed)
    return nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.GELU(), nn.Flatten(), nn.Linear(8 * 6 * 6, 300), nn.GELU(),
                         nn.Linear(300, 5))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from imageclassification_b200.ddp import DistributedDataParallel
    try:
        torch.set_num_threads(1)
        model = _model(100 + rank)                      # different init per rank: the broadcast must fix it
        ddp = DistributedDataParallel(model, bucket_cap_mb=0.005)      # tiny buckets -> several collectives
        ref = _model(100)
        for p, q in zip(model.parameters(), ref.parameters()):
            assert torch.equal(p, q)
        # every parameter has its own slot; slots start on 16-byte boundaries, so the arena may be a few words larger
        assert sum(b.numel for b in ddp.buckets) == ddp.arena.numel() >= sum(p.numel() for p in model.parameters())
        assert ddp.arena.numel() - sum(p.numel() for p in model.parameters()) < 4 * len(list(model.parameters()))
        assert len(ddp.buckets) > 2
        g = torch.Generator().manual_seed(7)
        X = torch.randn(2 * world, 3, 6, 6, generator=g)
        Y = torch.randn(2 * world, 5, generator=g)
        xs, ys = X[2 * rank:2 * rank + 2], Y[2 * rank:2 * rank + 2]
        opt = torch.optim.SGD(model.parameters(), lr=0.1)
        opt_ref = torch.optim.SGD(ref.parameters(), lr=0.1)
        for it in range(3):
            opt.zero_grad()                             # set_to_none: the hook must re-home fresh grads into the arena
            ((ddp(xs) - ys) ** 2).mean().backward()
            opt_ref.zero_grad()
            ((ref(X) - Y) ** 2).mean().backward()       # single process, concatenated batch
            for p, q in zip(model.parameters(), ref.parameters()):
                assert torch.allclose(p.grad, q.grad, rtol=1e-5, atol=1e-7), it
                assert p.grad.untyped_storage().data_ptr() == ddp.arena.untyped_storage().data_ptr()
            opt.step()
            opt_ref.step()
        # gradient accumulation over 2 micro-steps without zero_grad: all-reduce on each (the reference has no no_sync)
        opt.zero_grad()
        opt_ref.zero_grad()
        for _ in range(2):
            (((ddp(xs) - ys) ** 2).mean() / 2).backward()
            (((ref(X) - Y) ** 2).mean() / 2).backward()
        for p, q in zip(model.parameters(), ref.parameters()):
            assert torch.allclose(p.grad, q.grad, rtol=1e-5, atol=1e-7)
        ret[rank] = "ok"
    except Exception as e:  # noqa: BLE001
        import traceback
        ret[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


def test_ddp_world2_gloo():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}, dict(ret)


def test_ddp_requires_process_group():
    from imageclassification_b200.ddp import DistributedDataParallel
    with pytest.raises(RuntimeError):
        DistributedDataParallel(nn.Linear(2, 2))
