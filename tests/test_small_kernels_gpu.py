"""a9 SoftTargetCrossEntropy, a10 ModelEmaV3, a11 mixup_target: CUDA path (through the C-ABI) vs the oracle.
Bars (BASELINE.json north_star): EMA update and mixup label mixing BIT-EXACT in fp32; loss fp32 <= 1e-4 relative."""
import math

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import ema as o_ema, loss as o_loss, mixup as o_mix

pytestmark = pytest.mark.gpu
DEV = "cuda"


class _Bag(nn.Module):
    """A module whose state dict has ragged tensor sizes (1 element .. > one 8192-element chunk) and an int buffer."""

    def __init__(self, sizes, seed):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.ps = nn.ParameterList([nn.Parameter(torch.randn(s, generator=g)) for s in sizes])
        self.register_buffer("steps", torch.tensor(3, dtype=torch.int64))
        self.register_buffer("running", torch.randn(33, generator=g))


SIZES = [1, 3, 7, 96, 8191, 8192, 8193, 4 * 8192 + 5, (96, 1, 7, 7), (384, 96), 100003]


def _ema_pair(seed=0):
    import imageclassification_b200 as P
    cpu_model = _Bag(SIZES, seed)
    gpu_model = _Bag(SIZES, seed).to(DEV)
    o = o_ema.ModelEmaV3(cpu_model, decay=0.9995)
    e = P.ModelEmaV3(gpu_model, decay=0.9995, device=torch.device(DEV))
    return cpu_model, gpu_model, o, e


def _assert_bit_equal(a, b):
    a, b = a.detach().cpu().contiguous(), b.detach().cpu().contiguous()
    assert a.dtype == b.dtype and a.shape == b.shape
    assert torch.equal(a.view(torch.int32) if a.dtype == torch.float32 else a,
                       b.view(torch.int32) if b.dtype == torch.float32 else b)


@pytest.mark.parametrize("steps", [1, 2, 50])
def test_ema_bit_exact(steps):
    cpu_model, gpu_model, o, e = _ema_pair()
    g = torch.Generator().manual_seed(1)
    for s in range(steps):
        with torch.no_grad():
            for pc, pg in zip(cpu_model.parameters(), gpu_model.parameters()):
                d = torch.randn(pc.shape, generator=g) * 0.1
                pc.add_(d)
                pg.add_(d.to(DEV))
            cpu_model.steps += 1
            gpu_model.steps += 1
        o.update(cpu_model)
        e.update(gpu_model)
    for (k, a), (_, b) in zip(e.module.state_dict().items(), o.module.state_dict().items()):
        _assert_bit_equal(a, b)


def test_ema_special_values_and_set():
    cpu_model, gpu_model, o, e = _ema_pair(3)
    special = torch.tensor([0.0, -0.0, 1e-45, -1e-45, 1e-38, 3.4e38, -3.4e38, float("inf"), float("-inf"), float("nan"),
                            1.0, -1.0, 1e-20, 123456.789])
    with torch.no_grad():
        cpu_model.ps[4][:14] = special
        gpu_model.ps[4][:14] = special.to(DEV)
        o.module.ps[4][:14] = special.flip(0)
        e.module.ps[4][:14] = special.flip(0).to(DEV)
    o.update(cpu_model)
    e.update(gpu_model)
    a, b = e.module.ps[4].detach().cpu(), o.module.ps[4].detach()
    nan_a, nan_b = torch.isnan(a), torch.isnan(b)
    assert torch.equal(nan_a, nan_b)
    assert torch.equal(a[~nan_a].view(torch.int32), b[~nan_b].view(torch.int32))
    e.set(gpu_model)
    for a, b in zip(e.module.state_dict().values(), gpu_model.state_dict().values()):
        _assert_bit_equal(a, b)


def test_ema_through_wrapper_and_decay_schedule():
    import imageclassification_b200 as P
    m = _Bag([10, 9000], 5).to(DEV)
    wrapper = nn.Sequential()
    wrapper.module = m      # DDP-like: state_dict values are zipped, keys ignored (engine.py:68 passes the DDP wrapper)
    e = P.ModelEmaV3(m, decay=0.9, device=torch.device(DEV))
    assert not e.module.training
    before = [v.clone() for v in e.module.state_dict().values()]
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(2.0)
    e.update(wrapper)
    for b, a, p in zip(before, e.module.state_dict().values(), m.state_dict().values()):
        if a.is_floating_point():
            exp = torch.lerp(b, p, 1.0 - 0.9)
            assert torch.equal(a, exp)
    assert e.get_decay() == 0.9 and e.get_decay(0) == 0.0


@pytest.mark.parametrize("B,K", [(8, 2), (256, 1000), (64, 1000), (2, 1), (6, 37)])
@pytest.mark.parametrize("smoothing", [0.0, 0.1])
def test_mixup_target_bit_exact(B, K, smoothing):
    import imageclassification_b200 as P
    g = torch.Generator().manual_seed(B * 1000 + K)
    t = torch.randint(0, K, (B,), generator=g)
    rs = np.random.RandomState(7)
    lams = [0.0, 1.0, 0.5, 1.0 / 3.0] + [float(rs.beta(0.8, 0.8)) for _ in range(6)]
    for lam in lams:
        exp = o_mix.mixup_target(t, K, lam, smoothing)
        got = P.mixup_target(t.to(DEV), K, lam, smoothing)
        assert got.dtype == torch.float32 and got.shape == (B, K)
        assert torch.equal(got.cpu().view(torch.int32), exp.view(torch.int32)), f"lam={lam}"


def test_mixup_call_matches_oracle():
    import imageclassification_b200 as P
    for kw in [dict(mixup_alpha=0.8, cutmix_alpha=0.0), dict(mixup_alpha=0.8, cutmix_alpha=1.0),
               dict(mixup_alpha=0.0, cutmix_alpha=1.0), dict(mixup_alpha=0.0, cutmix_alpha=0.0, cutmix_minmax=(0.2, 0.8)),
               dict(mixup_alpha=0.8, cutmix_alpha=1.0, prob=0.5)]:
        o = o_mix.Mixup(label_smoothing=0.1, num_classes=10, **kw)
        p = P.Mixup(label_smoothing=0.1, num_classes=10, **kw)
        g = torch.Generator().manual_seed(11)
        for it in range(4):
            x = torch.randn(8, 3, 32, 32, generator=g)
            t = torch.randint(0, 10, (8,), generator=g)
            np.random.seed(100 + it)
            xo, to = o(x.clone(), t)
            np.random.seed(100 + it)
            xg, tg = p(x.clone().to(DEV), t.to(DEV))
            assert torch.equal(tg.cpu().view(torch.int32), to.view(torch.int32))
            # image mixing: same op sequence as timm; CUDA contracts mul+add differently from the CPU -> 1 ulp
            assert torch.allclose(xg.cpu(), xo, rtol=0, atol=1e-6)
    with pytest.raises(ValueError):
        P.Mixup()(torch.zeros(3, 3, 4, 4, device=DEV), torch.zeros(3, dtype=torch.long, device=DEV))


@pytest.mark.parametrize("shape", [(8, 3, 32, 32), (2, 3, 224, 224), (6, 1, 5, 7), (4, 3, 9, 9), (64, 3, 64, 64), (5, 2, 4, 4)])
def test_mixup_batch_bit_exact(shape):
    """cnx_mixup_batch (one in-place pass) vs timm's own tensor expressions run by ATen on the same device: bit-exact,
    for the blend (incl. lam = 0 / 1 / Beta draws, odd L, odd B through the C-ABI) and for the cutmix box swap; the optional
    un-mixed copy is exact too."""
    from imageclassification_b200 import ops
    g = torch.Generator().manual_seed(sum(shape))
    x0 = torch.randn(*shape, generator=g).to(DEV)
    x0.view(-1)[:4] = torch.tensor([0.0, -0.0, 1e-41, float("inf")], device=DEV)
    rs = np.random.RandomState(3)
    B, C, H, W = shape
    for lam in [0.0, 1.0, 0.5, 1.0 / 3.0] + [float(rs.beta(0.8, 0.8)) for _ in range(4)]:
        ref = x0.clone()
        flipped = ref.flip(0).mul_(1.0 - lam)
        ref.mul_(lam).add_(flipped)
        got, orig = x0.clone(), torch.full_like(x0, 7.0)
        ops.mixup_batch(got, lam, original_out=orig)
        _assert_bit_equal(got.nan_to_num(1e30), ref.nan_to_num(1e30))
        _assert_bit_equal(orig, x0)
        got2 = x0.clone()
        ops.mixup_batch(got2, lam)
        _assert_bit_equal(got2.nan_to_num(1e30), ref.nan_to_num(1e30))
    if B % 2 == 0:
        for box in [(0, H, 0, W), (1, H - 1, 2, W - 1), (0, 0, 0, 0), (H // 2, H // 2 + 1, 0, 1)]:
            yl, yh, xl, xh = box
            ref = x0.clone()
            ref[:, :, yl:yh, xl:xh] = ref.flip(0)[:, :, yl:yh, xl:xh]
            got, orig = x0.clone(), torch.empty_like(x0)
            ops.mixup_batch(got, 0.3, box=box, original_out=orig)
            _assert_bit_equal(got, ref)
            _assert_bit_equal(orig, x0)
    with pytest.raises(TypeError):
        ops.mixup_batch(x0.half(), 0.5)
    with pytest.raises(RuntimeError):
        ops.mixup_batch(x0.clone(), 0.5, box=(0, H + 1, 0, 1))


@pytest.mark.parametrize("B,K", [(8, 2), (256, 1000), (64, 1000), (512, 1000), (3, 5000), (1, 1)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_soft_target_ce(B, K, dtype):
    import imageclassification_b200 as P
    g = torch.Generator().manual_seed(B + K)
    x = (torch.randn(B, K, generator=g) * 3).to(dtype)
    t = o_mix.mixup_target(torch.randint(0, K, (B,), generator=g), K, 0.37, 0.1)
    xo = x.float().clone().requires_grad_(True)        # oracle: CUDA autocast runs log_softmax in fp32
    lo = o_loss.SoftTargetCrossEntropy()(xo, t)
    (lo * 1.7).backward()
    xg = x.to(DEV).requires_grad_(True)
    lg = P.SoftTargetCrossEntropy()(xg, t.to(DEV))
    assert lg.dim() == 0 and lg.dtype == torch.float32
    (lg * 1.7).backward()
    assert abs(lg.item() - lo.item()) <= 1e-5 * max(1.0, abs(lo.item()))      # fp32 bar: 1e-4 relative
    tol = 1e-5 if dtype == torch.float32 else 1e-2                            # bf16 bar: 2e-2 (grad is stored in bf16)
    assert xg.grad.dtype == dtype
    den = xo.grad.abs().max().item() + 1e-30
    assert (xg.grad.float().cpu() - xo.grad).abs().max().item() / den <= tol


def test_soft_target_ce_known_answers():
    import imageclassification_b200 as P
    crit = P.SoftTargetCrossEntropy()
    K = 1000
    x = torch.zeros(4, K, device=DEV)
    t = torch.full((4, K), 1.0 / K, device=DEV)
    assert abs(crit(x, t).item() - math.log(K)) < 1e-5                        # uniform logits -> ln K
    x = torch.randn(16, 10, device=DEV)
    lab = torch.randint(0, 10, (16,), device=DEV)
    oh = torch.nn.functional.one_hot(lab, 10).float()
    assert abs(crit(x, oh).item() - torch.nn.functional.cross_entropy(x, lab).item()) < 1e-5
    t2 = torch.rand(16, 10, device=DEV) * 3.0                                 # targets that do not sum to 1
    ref = o_loss.SoftTargetCrossEntropy()(x.cpu(), t2.cpu()).item()
    assert abs(crit(x, t2).item() - ref) < 1e-5 * abs(ref)
    x[3, 4] = float("nan")                                                    # engine.py:56-59 relies on NaN propagating
    assert not math.isfinite(crit(x, oh).item())
    # loss /= update_freq ; loss.backward()  (engine.py:71-72)
    x = torch.randn(8, 5, device=DEV, requires_grad=True)
    loss = crit(x, torch.softmax(torch.randn(8, 5, device=DEV), -1))
    loss /= 4
    loss.backward()
    assert x.grad is not None and torch.isfinite(x.grad).all()
    with pytest.raises(RuntimeError):
        crit(torch.zeros(2, 3), torch.zeros(2, 3))                            # no CPU fallback
