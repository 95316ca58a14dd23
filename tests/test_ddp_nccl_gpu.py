"""Multi-rank parity over NCCL: world_size-2 run of imageclassification_b200.ddp.DistributedDataParallel
on two GPUs of one box against single-process training on the concatenated batch, with the real ConvNeXt-T and the libcnx
kernels (fp32 bar 1e-4 on gradients and updated parameters; bf16 autocast 2e-2).  Skipped below 2 GPUs (the CPU suite covers
the same host logic over gloo in tests/test_ddp_cpu.py)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _data(world, per_rank, K):
    g = torch.Generator().manual_seed(17)
    x = torch.randn(world * per_rank, 3, 64, 64, generator=g)
    t = torch.rand(world * per_rank, K, generator=g).softmax(-1)          # soft targets, as mixup hands them to the criterion
    return x, t


def _train(model, x, t, amp, steps, P):
    crit = P.SoftTargetCrossEntropy()
    opt = torch.optim.SGD(model.parameters(), lr=0.05)
    grads = None
    for _ in range(steps):
        opt.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            loss = crit(model(x), t)
        loss.backward()
        grads = [p.grad.detach().clone() for p in model.parameters()]
        opt.step()
    return grads


def _worker(rank, world, port, amp, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    try:
        import torch.distributed as dist
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        import imageclassification_b200 as P
        from cabi import max_rel
        from imageclassification_b200.ddp import DistributedDataParallel
        K, per = 8, 4
        x, t = _data(world, per, K)
        torch.manual_seed(100 + rank)                                     # different init per rank: the broadcast must fix it
        model = P.create_model("convnext_tiny", num_classes=K, ls_init_value=1.0).to(dev)
        ddp = DistributedDataParallel(model, device_ids=[rank], bucket_cap_mb=4.0)
        assert len(ddp.buckets) > 3
        torch.manual_seed(100)
        ref = P.create_model("convnext_tiny", num_classes=K, ls_init_value=1.0).to(dev)
        for p, q in zip(model.parameters(), ref.parameters()):
            assert torch.equal(p, q)
        sl = slice(rank * per, (rank + 1) * per)
        g_ddp = _train(ddp, x[sl].to(dev), t[sl].to(dev), amp, 2, P)
        g_ref = _train(ref, x.to(dev), t.to(dev), amp, 2, P)              # single process, concatenated batch
        tol = 2e-2 if amp else 1e-4
        worst = 0.0
        for (n, p), a, b, q in zip(model.named_parameters(), g_ddp, g_ref, ref.parameters()):
            assert p.grad.untyped_storage().data_ptr() == ddp.arena.untyped_storage().data_ptr(), n
            worst = max(worst, max_rel(a, b))
            assert max_rel(a, b) <= tol, (n, max_rel(a, b))
            assert max_rel(p.detach(), q.detach()) <= tol, n
        # every rank holds the same averaged gradients
        flat = torch.cat([g.flatten() for g in g_ddp])
        other = flat.clone()
        dist.broadcast(other, src=0)
        assert torch.equal(flat, other)
        ret[rank] = f"ok {worst:.2e}"
        dist.destroy_process_group()
    except Exception:  # noqa: BLE001
        import traceback
        ret[rank] = traceback.format_exc()


@pytest.mark.parametrize("amp", [False, True], ids=["fp32", "bf16"])
def test_ddp_world2_nccl_matches_single_process(amp):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs on one box")
    import torch.multiprocessing as mp
    port = 29700 + (os.getpid() % 1000) + (1 if amp else 0)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, amp, ret), nprocs=2, join=True)
    assert all(str(ret.get(r, "")).startswith("ok") for r in (0, 1)), dict(ret)
