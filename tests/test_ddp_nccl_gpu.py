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


@pytest.mark.parametrize("amp", [False, True], ids=["fp32", "bf16"])
def test_direct_arena_gradients_single_rank(amp):
    """One-rank process group on one GPU: the libcnx backward functions write the Block / stem / head parameter gradients
    straight into the reducer's arena (no autograd accumulation, no copy) — same values as the unwrapped model, with
    zero_grad(set_to_none) between steps, gradient accumulation over two micro-steps, and `direct_grads=False` as the control."""
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import imageclassification_b200 as P
    from cabi import max_rel
    from imageclassification_b200.ddp import DistributedDataParallel
    own_group = not dist.is_initialized()
    if own_group:
        dist.init_process_group("gloo", store=dist.HashStore(), rank=0, world_size=1)
    try:
        _direct_arena_body(amp, P, DistributedDataParallel, max_rel)
    finally:
        if own_group:
            dist.destroy_process_group()


def _direct_arena_body(amp, P, DistributedDataParallel, max_rel):
    dev = torch.device("cuda")
    K = 8
    x, t = _data(1, 4, K)
    x, t = x.to(dev), t.to(dev)
    crit = P.SoftTargetCrossEntropy()

    def make():
        torch.manual_seed(5)
        return P.create_model("convnext_tiny", num_classes=K, ls_init_value=1.0, drop_path_rate=0.05).to(dev)

    def run(model, net, accum):
        opt = torch.optim.SGD(model.parameters(), lr=0.05)
        torch.manual_seed(9)                                    # drop-path masks
        for it in range(2):
            opt.zero_grad()
            for _ in range(accum):
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                    loss = crit(net(x), t) / accum
                loss.backward()
            grads = [p.grad.detach().clone() for p in model.parameters()]
            opt.step()
        return grads

    for accum in (1, 2):
        ref = make()
        g_ref = run(ref, ref, accum)
        for direct in (True, False):
            m = make()
            ddp = DistributedDataParallel(m, device_ids=[0], bucket_cap_mb=4.0, direct_grads=direct)
            g = run(m, ddp, accum)
            for (n, p), a, b, q in zip(m.named_parameters(), g, g_ref, ref.parameters()):
                assert p.grad.untyped_storage().data_ptr() == ddp.arena.untyped_storage().data_ptr(), n
                assert p.grad.data_ptr() == ddp._slot[p][1].data_ptr(), n
                if accum == 1:
                    assert torch.equal(a, b), (n, direct, max_rel(a, b))
                    assert torch.equal(p.detach(), q.detach()), n
                else:                       # accumulating kernels add the second micro-step's partial sums to the first's
                    assert max_rel(a, b) <= 2e-5, (n, direct, max_rel(a, b))     # result (another association than autograd's add)
                    assert max_rel(p.detach(), q.detach()) <= 2e-5, n
