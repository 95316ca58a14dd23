"""engine.train_one_epoch end to end (BASELINE configs 1 and 3): this package's objects and step
loop on the B200 vs
  * the golden numbers the REFERENCE'S OWN engine.py produced in the build container (tests/golden/engine_step.npz,
    case "b": no drop-path, so no device-RNG dependence), and
  * the oracle objects + oracle step loop run on the same device with the same seeds (drop-path, cutmix, gradient
    accumulation, bf16 autocast, fused AdamW+EMA optimizer).
Bars: fp32 <= 1e-4 relative on loss / parameter norms, bf16 <= 2e-2 (BASELINE.json north_star); EMA follows parameters."""
import os

import numpy as np
import pytest
import torch

import imageclassification_b200 as P
from imageclassification_b200 import engine as PE
from imageclassification_b200 import optim as PO
from oracle import convnext as OC, ema as OE, engine as OEng, loss as OL, mixup as OM

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _adamw(model, lr=1e-3, wd=5e-4):
    return torch.optim.AdamW([{"params": list(model.parameters()), "weight_decay": wd}], lr=lr, weight_decay=0.0)


def _norms(params):
    return np.array([p.detach().double().norm().item() for p in params])


@pytest.mark.parametrize("x3_train", [False, True], ids=["fp32_cuda_cores", "fp32_tensor_cores_x3"])
def test_engine_matches_reference_golden_config1(x3_train, monkeypatch):
    """BASELINE config 1 (ConvNeXt-T fp32, batch 8, 2 classes, mixup 0.8, smoothing 0.1, AdamW, EMA 0.9995), two iterations:
    same seeds as tests/golden/make_golden.py case "b" => the loss, accuracy, parameters and EMA the reference's engine gave.
    Run with the fp32 CUDA-core GEMMs (errors ~1e-7: every parameter norm to 1e-4) and with the default split-operand
    tensor-core GEMMs (gradients to ~1e-5, inside the 1e-4 gradient bar of tests/test_block_gpu.py; after two Adam steps a
    zero-initialised bias element whose gradient is itself ~1e-5 of the tensor's largest can land elsewhere, because Adam moves
    every element by ~lr whatever its gradient's size — hence the absolute floor for those 1e-2-norm tensors)."""
    from imageclassification_b200 import ops
    monkeypatch.setattr(ops, "X3_TRAIN", x3_train)
    z = np.load(os.path.join(GOLD, "engine_step.npz"))
    torch.manual_seed(88)
    np.random.seed(88)
    o = OC.create_model("convnext_tiny", num_classes=2, drop_path_rate=0.0, ls_init_value=1.0)      # CPU init = golden's init
    data = [(torch.randn(8, 3, 64, 64), torch.randint(0, 2, (8,))) for _ in range(2)]
    model = P.create_model("convnext_tiny", num_classes=2, drop_path_rate=0.0, ls_init_value=1.0)
    model.load_state_dict(o.state_dict())
    model.to(DEV)
    ema = P.ModelEmaV3(model, decay=0.9995, device=DEV)
    mix = P.Mixup(mixup_alpha=0.8, cutmix_alpha=0.0, label_smoothing=0.1, num_classes=2)
    stats = PE.train_one_epoch(model, P.SoftTargetCrossEntropy(), data, _adamw(model), DEV, 0, None, None, ema, mix,
                               num_training_steps_per_epoch=2, update_freq=1, use_amp=False, num_classes=2, verbose=False)
    assert abs(stats["loss"] - float(z["b.loss"])) <= 1e-4 * abs(float(z["b.loss"]))
    assert stats["class_acc"] == float(z["b.class_acc"])
    floor = 3e-4 if x3_train else 1e-6
    np.testing.assert_allclose(_norms(model.parameters()), z["b.param_norms"], rtol=1e-4, atol=floor)
    np.testing.assert_allclose(_norms(ema.module.parameters()), z["b.ema_norms"], rtol=1e-5, atol=1e-6 if x3_train else 1e-8)
    ps = np.array([p.detach().double().sum().item() for p in model.parameters()])
    np.testing.assert_allclose(ps, z["b.param_sums"], rtol=0, atol=(4e-3 if x3_train else 2e-3) * np.abs(z["b.param_norms"]).max())


CASES = [
    # tag, amp, img, batch, classes, dpr, gamma, cutmix, update_freq, fused optimizer
    ("fp32_dp", False, 64, 8, 2, 0.05, 1.0, 0.0, 1, False),
    ("fp32_cutmix_accum", False, 64, 8, 5, 0.0, 1.0, 1.0, 2, False),
    ("fp32_fused_opt", False, 64, 8, 2, 0.0, 1.0, 0.0, 1, True),
    ("bf16", True, 96, 16, 10, 0.05, 1.0, 1.0, 1, True),
    ("bf16_tiny_gamma", True, 64, 8, 2, 0.0, 1e-6, 0.0, 2, False),
]


@pytest.mark.parametrize("tag,amp,img,batch,K,dpr,gamma,cutmix,uf,fused", CASES, ids=[c[0] for c in CASES])
def test_engine_matches_oracle_engine(tag, amp, img, batch, K, dpr, gamma, cutmix, uf, fused):
    torch.manual_seed(11)
    o = OC.create_model("convnext_tiny", num_classes=K, drop_path_rate=dpr, ls_init_value=gamma).to(DEV)
    p = P.create_model("convnext_tiny", num_classes=K, drop_path_rate=dpr, ls_init_value=gamma).to(DEV)
    p.load_state_dict(o.state_dict())
    g = torch.Generator().manual_seed(5)
    n_it = 2 * uf
    data = [(torch.randn(batch, 3, img, img, generator=g), torch.randint(0, K, (batch,), generator=g)) for _ in range(n_it)]
    out = []
    for eng, model, crit, mixc, emac in ((OEng, o, OL.SoftTargetCrossEntropy(), OM.Mixup, OE.ModelEmaV3),
                                         (PE, p, P.SoftTargetCrossEntropy(), P.Mixup, P.ModelEmaV3)):
        torch.manual_seed(3)                   # drop-path masks
        np.random.seed(3)                      # mixup lambda / cutmix box
        ema = emac(model, decay=0.9995, device=DEV)
        mix = mixc(mixup_alpha=0.8, cutmix_alpha=cutmix, label_smoothing=0.1, num_classes=K)
        if fused and eng is PE:
            opt = PO.AdamW([{"params": list(model.parameters()), "weight_decay": 5e-4}], lr=1e-3, weight_decay=0.0)
            opt.fuse_ema(ema, model)
        else:
            opt = _adamw(model)
        kw = {"verbose": False} if eng is PE else {}
        stats = eng.train_one_epoch(model, crit, [(a.clone(), b.clone()) for a, b in data], opt, DEV, 0, None, None, ema, mix,
                                    num_training_steps_per_epoch=2, update_freq=uf, use_amp=amp, num_classes=K, **kw)
        out.append((stats, _norms(model.parameters()), _norms(ema.module.parameters()),
                    torch.cat([q.detach().flatten() for q in model.parameters()]).double()))
    (so, no, eo, fo), (sp, npar, ep, fp) = out
    tol = 2e-2 if amp else 1e-4
    assert abs(sp["loss"] - so["loss"]) <= tol * abs(so["loss"]), (sp, so)
    # zero-initialised biases are pure Adam updates (~lr per element per step, direction set by noisy bf16 gradients): absolute floor
    np.testing.assert_allclose(npar, no, rtol=tol, atol=1e-3 if amp else 1e-6)
    np.testing.assert_allclose(ep, eo, rtol=tol, atol=1e-6 if amp else 1e-8)      # the EMA follows the parameters
    if not amp:
        # Adam's first steps move every weight by ~lr*sign(g): compare the parameter vectors themselves (gradient-sign flips on
        # near-zero gradients are the only admissible differences; they are bounded by 2*lr per element)
        d = (fp - fo).abs()
        assert d.max().item() <= 2 * 2e-3 + 1e-6
        assert (d > 1e-5).double().mean().item() <= 1e-2
        assert sp["class_acc"] == so["class_acc"]


def test_prefetched_batches_give_identical_results():
    """DevicePrefetcher only moves the H2D copy to a side stream: bit-identical parameters with and without it, for pinned,
    pageable and device-resident batches."""
    g = torch.Generator().manual_seed(9)
    data = [(torch.randn(8, 3, 64, 64, generator=g), torch.randint(0, 3, (8,), generator=g)) for _ in range(3)]
    variants = {"pageable": lambda: data, "pinned": lambda: [(a.pin_memory(), b.pin_memory()) for a, b in data],
                "device": lambda: [(a.to(DEV), b.to(DEV)) for a, b in data]}     # mixup mutates device batches in place: fresh copies
    res = {}
    for name, make in variants.items():
        for pf in (False, True):
            batches = make()
            torch.manual_seed(21)
            np.random.seed(21)
            m = P.create_model("convnext_tiny", num_classes=3, drop_path_rate=0.05, ls_init_value=1.0).to(DEV)
            mix = P.Mixup(mixup_alpha=0.8, cutmix_alpha=1.0, label_smoothing=0.1, num_classes=3)
            st = PE.train_one_epoch(m, P.SoftTargetCrossEntropy(), batches, _adamw(m), DEV, 0, None, None, None, mix,
                                    use_amp=True, num_classes=3, verbose=False, prefetch=pf)
            res[(name, pf)] = (st["loss"], torch.cat([q.detach().flatten() for q in m.parameters()]))
    l0, p0 = res[("pageable", False)]
    for k, (l, p) in res.items():
        assert l == l0 and torch.equal(p, p0), k
    assert len(PE.DevicePrefetcher(data, DEV)) == 3


def test_evaluate_matches_reference_golden_and_oracle():
    """engine.evaluate: same keys and values as the reference's own evaluate on the golden case (fp32),
    and as the oracle's evaluate on the device under bf16 autocast (loss within the bf16 bar; counts from the same argmax)."""
    z = np.load(os.path.join(GOLD, "evaluate.npz"))
    gold = dict(zip([str(k) for k in z["keys"]], z["values"].tolist()))
    torch.manual_seed(5)
    o = OC.create_model("convnext_tiny", num_classes=3, ls_init_value=1.0)
    with torch.no_grad():
        o.head.fc.weight.normal_(0, 0.5)
    data = [(torch.randn(b, 3, 64, 64), torch.randint(0, 3, (b,))) for b in (8, 8, 5)]
    p = P.create_model("convnext_tiny", num_classes=3, ls_init_value=1.0)
    p.load_state_dict(o.state_dict())
    p.to(DEV)
    p.train()                                                    # evaluate() must switch to eval mode itself
    stats = PE.evaluate(data, p, DEV, 3, use_amp=False, verbose=False)
    assert not p.training and sorted(stats) == sorted(gold)
    for k, v in gold.items():
        assert abs(stats[k] - v) <= 1e-4 * max(1.0, abs(v)), (k, stats[k], v)
    # a 3-tuple loader (engine.py:170-171 takes batch[0], batch[-1]) and no prefetch
    stats2 = PE.evaluate([(a, None, b) for a, b in data], p, DEV, 3, use_amp=False, verbose=False, prefetch=False)
    assert stats2 == stats
    o.to(DEV)
    so = OEng.evaluate(data, o, DEV, 3, use_amp=True)
    sp = PE.evaluate(data, p, DEV, 3, use_amp=True, verbose=False)
    assert abs(sp["loss"] - so["loss"]) <= 2e-2 * abs(so["loss"])
    assert abs(sp["acc1"] - so["acc1"]) <= 100.0 / 21 + 1e-6       # at most one near-tie argmax flip among 21 samples


def test_nested_prefetchers_do_not_share_buffers():
    """Two loaders iterated at the same time on one device (a validation pass inside a training loop): the second one gets its
    own stream and buffers, so neither sees the other's batches."""
    g = torch.Generator().manual_seed(4)
    a = [(torch.randn(4, 3, 8, 8, generator=g), torch.full((4,), i)) for i in range(5)]
    b = [(torch.randn(4, 3, 8, 8, generator=g), torch.full((4,), 100 + i)) for i in range(3)]
    seen = []
    for xa, ta in PE.DevicePrefetcher(a, DEV):
        inner = [(xb.clone(), tb.clone()) for xb, tb in PE.DevicePrefetcher(b, DEV)]
        assert [int(t[0]) for _, t in inner] == [100, 101, 102]
        assert all(torch.equal(x.cpu(), b[i][0]) for i, (x, _) in enumerate(inner))
        seen.append((xa.clone(), int(ta[0])))
    assert [t for _, t in seen] == [0, 1, 2, 3, 4]
    assert all(torch.equal(x.cpu(), a[i][0]) for i, (x, _) in enumerate(seen))
