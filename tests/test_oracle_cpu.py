"""Pins the oracle (CPU, no GPU needed):
  * against golden vectors produced by the REFERENCE'S OWN code (tests/golden/make_golden.py): the in-tree Block /
    LayerNorm of semantic_segmentation/backbone/convnext.py and two iterations of the reference engine.train_one_epoch;
  * against analytic known answers for the timm-only pieces (SoftTargetCrossEntropy, ModelEmaV3, mixup_target)."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import convnext as OC, ema as OE, engine as OEng, loss as OL, mixup as OM

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_block(C, H, gamma_init=1.0):
    z = np.load(os.path.join(GOLD, f"block_C{C}_H{H}.npz"))
    blk = OC.ConvNeXtBlock(C, ls_init_value=gamma_init)
    blk.load_state_dict({k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("p.")})
    return z, blk


@pytest.mark.parametrize("C,H", [(32, 9), (96, 7), (64, 12)])
def test_oracle_block_matches_reference_block(C, H):
    z, blk = load_block(C, H)
    x = torch.from_numpy(z["x"]).requires_grad_(True)
    y = blk(x)
    y.backward(torch.from_numpy(z["dout"]))
    assert torch.equal(y.detach(), torch.from_numpy(z["y"]))                 # same ops, same order: bit-identical
    assert torch.allclose(x.grad, torch.from_numpy(z["dx"]), rtol=1e-6, atol=1e-7)
    for n, p in blk.named_parameters():
        g = torch.from_numpy(z["g." + n])
        assert torch.allclose(p.grad, g, rtol=1e-5, atol=1e-6 * g.abs().max().item()), n


def test_oracle_layernorm_matches_reference_both_formats():
    z = np.load(os.path.join(GOLD, "ln_channels_first.npz"))
    x = torch.from_numpy(z["x"])
    C = x.shape[-1]
    cl, cf = OC.LayerNorm(C), OC.LayerNorm2d(C)
    with torch.no_grad():
        for m in (cl, cf):
            m.weight.copy_(torch.from_numpy(z["w"]))
            m.bias.copy_(torch.from_numpy(z["b"]))
        assert torch.equal(cl(x), torch.from_numpy(z["y_cl"]))
        assert torch.allclose(cf(x.permute(0, 3, 1, 2)), torch.from_numpy(z["y_cf"]), rtol=1e-5, atol=2e-6)
        # float64 biased-variance formula
        xd = x.double()
        ref = (xd - xd.mean(-1, keepdim=True)) / torch.sqrt(xd.var(-1, unbiased=False, keepdim=True) + 1e-6)
        ref = ref * torch.from_numpy(z["w"]).double() + torch.from_numpy(z["b"]).double()
        assert (cl(x).double() - ref).abs().max() < 2e-6


@pytest.mark.parametrize("tag,img,gamma_init,dpr", [("b", 64, 1.0, 0.0), ("a", 224, 1e-6, 0.05)])
def test_oracle_engine_matches_reference_engine(tag, img, gamma_init, dpr):
    """oracle/engine.py restates engine.py:27-97; same seeds => same loss, accuracy, parameters and EMA as the reference's
    own engine.train_one_epoch produced in the build container."""
    z = np.load(os.path.join(GOLD, "engine_step.npz"))
    torch.set_num_threads(8)
    torch.manual_seed(88)
    np.random.seed(88)
    model = OC.create_model("convnext_tiny", num_classes=2, drop_path_rate=dpr, ls_init_value=gamma_init)
    ema = OE.ModelEmaV3(model, decay=0.9995, device=torch.device("cpu"))
    opt = torch.optim.AdamW([{"params": list(model.parameters()), "weight_decay": 5e-4}], lr=1e-3, weight_decay=0.0)
    mix = OM.Mixup(mixup_alpha=0.8, cutmix_alpha=0.0, label_smoothing=0.1, num_classes=2)
    data = [(torch.randn(8, 3, img, img), torch.randint(0, 2, (8,))) for _ in range(2)]
    stats = OEng.train_one_epoch(model, OL.SoftTargetCrossEntropy(), data, opt, "cpu", 0, None, None, ema, mix,
                                 num_training_steps_per_epoch=2, update_freq=1, use_amp=False, num_classes=2)
    assert abs(stats["loss"] - float(z[f"{tag}.loss"])) <= 1e-5 * abs(float(z[f"{tag}.loss"]))
    assert stats["class_acc"] == float(z[f"{tag}.class_acc"])
    pn = np.array([p.detach().double().norm().item() for p in model.parameters()])
    en = np.array([p.detach().double().norm().item() for p in ema.module.parameters()])
    ps = np.array([p.detach().double().sum().item() for p in model.parameters()])
    np.testing.assert_allclose(pn, z[f"{tag}.param_norms"], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(en, z[f"{tag}.ema_norms"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(ps, z[f"{tag}.param_sums"], rtol=0, atol=2e-3 * np.abs(z[f"{tag}.param_norms"]).max())


def _eval_case():
    torch.manual_seed(5)
    model = OC.create_model("convnext_tiny", num_classes=3, ls_init_value=1.0)
    with torch.no_grad():
        model.head.fc.weight.normal_(0, 0.5)
    data = [(torch.randn(b, 3, 64, 64), torch.randint(0, 3, (b,))) for b in (8, 8, 5)]
    return model, data


def test_oracle_evaluate_matches_reference_evaluate():
    """oracle/engine.py evaluate restates engine.py:145-225; the golden dict is what the reference's own evaluate returned."""
    z = np.load(os.path.join(GOLD, "evaluate.npz"))
    gold = dict(zip([str(k) for k in z["keys"]], z["values"].tolist()))
    torch.set_num_threads(8)
    model, data = _eval_case()
    stats = OEng.evaluate(data, model, "cpu", 3)
    assert sorted(stats) == sorted(gold)
    for k, v in gold.items():
        assert abs(stats[k] - v) <= 1e-5 * max(1.0, abs(v)), (k, stats[k], v)


def test_param_counts():
    # SURVEY.md §7 step 0 self-check
    assert sum(p.numel() for p in OC.create_model("convnext_tiny", num_classes=1000).parameters()) == 28_589_128
    assert sum(p.numel() for p in OC.create_model("convnext_tiny", num_classes=2).parameters()) == 27_821_666
    m = OC.create_model("convnext_tiny")
    assert len(m.state_dict()) == 182
    keys = set(m.state_dict())
    for k in ("stem.0.weight", "stem.1.bias", "stages.1.downsample.0.weight", "stages.1.downsample.1.weight",
              "stages.2.blocks.8.conv_dw.weight", "stages.3.blocks.2.mlp.fc2.bias", "stages.0.blocks.0.gamma",
              "head.norm.weight", "head.fc.bias"):
        assert k in keys
    assert m.state_dict()["stages.0.blocks.0.conv_dw.weight"].shape == (96, 1, 7, 7)


def test_soft_target_ce_known_answers():
    crit = OL.SoftTargetCrossEntropy()
    K = 1000
    assert abs(crit(torch.zeros(4, K), torch.full((4, K), 1.0 / K)).item() - math.log(K)) < 1e-5
    x = torch.randn(16, 10)
    lab = torch.randint(0, 10, (16,))
    oh = torch.nn.functional.one_hot(lab, 10).float()
    assert abs(crit(x, oh).item() - torch.nn.functional.cross_entropy(x, lab).item()) < 1e-6
    # gradient formula used by the CUDA backward: (softmax * sum_k t - t) / B
    x = torch.randn(5, 7, requires_grad=True)
    t = torch.rand(5, 7) * 2
    crit(x, t).backward()
    ref = (torch.softmax(x.detach(), -1) * t.sum(-1, keepdim=True) - t) / 5
    assert torch.allclose(x.grad, ref, atol=1e-6)


def test_ema_is_fma_lerp():
    """ModelEmaV3's update is ATen lerp with weight fp32(1-decay): e + w*(p-e) with a single rounding of the product-sum."""
    torch.manual_seed(0)
    lin = torch.nn.Linear(64, 33)
    ema = OE.ModelEmaV3(lin, decay=0.9995)
    e0 = [v.clone() for v in ema.module.state_dict().values()]
    with torch.no_grad():
        for p in lin.parameters():
            p.add_(torch.randn_like(p))
    ema.update(lin)
    w = np.float32(1.0 - 0.9995)
    for b, a, p in zip(e0, ema.module.state_dict().values(), lin.state_dict().values()):
        bd, pd = b.double().numpy(), p.double().numpy()
        diff = (p.numpy() - b.numpy()).astype(np.float64)            # fp32 subtraction, exact in fp64
        exact = (np.float64(w) * diff + bd).astype(np.float32)       # one rounding = fmaf
        assert np.array_equal(a.numpy(), exact)
    assert not ema.module.training and ema.get_decay(0) == 0.0


def test_mixup_target_known_answers():
    t = torch.tensor([0, 2, 1, 1])
    y = OM.mixup_target(t, 3, lam=1.0, smoothing=0.0)
    assert torch.equal(y, torch.nn.functional.one_hot(t, 3).float())
    y = OM.mixup_target(t, 3, lam=0.25, smoothing=0.0)
    assert torch.allclose(y[0], torch.tensor([0.25, 0.75, 0.0]))     # partner of sample 0 is sample 3 (flip)
    y = OM.mixup_target(t, 4, lam=0.5, smoothing=0.1)
    assert torch.allclose(y.sum(-1), torch.ones(4))
    np.random.seed(5)
    m = OM.Mixup(mixup_alpha=0.8, cutmix_alpha=1.0, num_classes=4)
    x = torch.randn(4, 3, 16, 16)
    x2, y2 = m(x.clone(), t)
    assert x2.shape == x.shape and y2.shape == (4, 4) and y2.dtype == torch.float32
    with pytest.raises(AssertionError):
        m(torch.randn(3, 3, 8, 8), torch.tensor([0, 1, 2]))
