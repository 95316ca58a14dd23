"""Checkpoint / resume compatibility on CPU — no kernels involved:

  * this package's save_model / auto_load_model / initialize_model (mirrors of utils.py:536-615 and val.py:14-28);
  * the REFERENCE'S OWN utils.save_model / utils.auto_load_model (and val.initialize_model when val.py's imports are
    available), imported unmodified from /root/reference through oracle/shim.py, operating on this package's modules:
    whole-module pickle, state-dict key/shape filter, EMA state, optimizer state.  Skipped where /root/reference is absent.
"""
import os
import types

import pytest
import torch

import imageclassification_b200 as P
from imageclassification_b200 import optim as PO, utils as PU

REF = "/root/reference"


def _perturbed(num_classes, seed):
    torch.manual_seed(seed)
    m = P.create_model("convnext_tiny", num_classes=num_classes, drop_path_rate=0.05)
    with torch.no_grad():
        for q in m.parameters():
            q.add_(0.01 * torch.randn_like(q))
    return m


def _args(**kw):
    base = dict(resume="", auto_resume=True, model_ema=True, start_epoch=0, eval=False, save_ckpt_num=3, save_ckpt_freq=1)
    base.update(kw)
    return types.SimpleNamespace(**base)


def _same(a, b):
    sa, sb = a.state_dict(), b.state_dict()
    return list(sa) == list(sb) and all(torch.equal(sa[k], sb[k]) for k in sa)


def test_package_checkpoint_round_trip(tmp_path):
    m = _perturbed(2, 1)
    ema = P.ModelEmaV3(m, decay=0.9995)
    with torch.no_grad():
        for q in ema.module.parameters():
            q.mul_(0.5)
    opt = PO.AdamW([{"params": list(m.parameters()), "weight_decay": 5e-4}], lr=1e-3, weight_decay=0.0)
    for q in m.parameters():                                      # optimizer state without a device: as a resumed run has it
        opt.state[q] = {"step": 7, "exp_avg": torch.full_like(q, 0.25), "exp_avg_sq": torch.full_like(q, 0.5)}
    scaler = P.NativeScaler()
    scaler.load_state_dict({"scale": 65536.0, "_growth_tracker": 3})          # a torch GradScaler state dict loads
    args = _args()
    path = PU.save_model(args, (224, 224), 11, m, opt, scaler, ema, 2, output_dir=tmp_path)
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert sorted(ck) == sorted(["model", "optimizer", "epoch", "scaler", "input_shape", "num_classes", "args", "model_ema"])
    assert isinstance(ck["model"], P.ConvNeXt) and ck["scaler"]["scale"] == 65536.0

    m2, args2 = _perturbed(2, 2), _args()
    ema2 = P.ModelEmaV3(m2, decay=0.9995)
    opt2 = PO.AdamW([{"params": list(m2.parameters()), "weight_decay": 5e-4}], lr=1e-3, weight_decay=0.0)
    PU.auto_load_model(args2, m2, opt2, P.NativeScaler(), ema2, output_dir=tmp_path)
    assert _same(m, m2) and _same(ema.module, ema2.module) and args2.start_epoch == 12
    assert all(int(s["step"]) == 7 and torch.equal(s["exp_avg"], torch.full_like(s["exp_avg"], 0.25)) for s in opt2.state.values())
    assert opt2._tables == {}                                     # device pointer tables are rebuilt after a load

    # a checkpoint with another number of classes: head.fc.* skipped by the key/shape filter, EMA re-seeded from the model,
    # optimizer state NOT loaded (utils.py:584-606)
    m3, args3 = _perturbed(5, 3), _args()
    ema3 = P.ModelEmaV3(m3, decay=0.9995)
    opt3 = PO.AdamW([{"params": list(m3.parameters()), "weight_decay": 5e-4}], lr=1e-3, weight_decay=0.0)
    fc_before = m3.head.fc.weight.detach().clone()
    PU.auto_load_model(args3, m3, opt3, P.NativeScaler(), ema3, output_dir=tmp_path)
    assert torch.equal(m3.head.fc.weight, fc_before) and torch.equal(m3.stem[0].weight, m.stem[0].weight)
    assert _same(ema3.module, m3) and len(opt3.state) == 0 and args3.start_epoch == 0

    net, k = PU.initialize_model(str(path), True, "cpu")
    assert k == 2 and not net.training and _same(net, ema.module)
    net, _ = PU.initialize_model(str(path), False, "cpu")
    assert _same(net, m)
    kept, skipped = PU.filter_state_dict(m.state_dict(), m3.state_dict())
    assert skipped == 2 and "head.fc.weight" not in kept and len(kept) == len(m.state_dict()) - 2


@pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference is only present in the build container")
def test_reference_utils_save_and_resume_product_modules(tmp_path, monkeypatch):
    """/root/reference/utils.py:536-615 unmodified on this package's model / EMA / optimizer / scaler objects."""
    from oracle import shim
    shim.import_reference_engine()                                # installs the timm / tensorboardX stand-ins, loads utils.py
    import sys
    rutils = sys.modules["utils"]
    assert rutils.__file__.startswith(REF)
    monkeypatch.chdir(tmp_path)
    os.makedirs("train_cls/output")
    m = _perturbed(2, 4)
    ema = P.ModelEmaV3(m, decay=0.9995)
    opt = PO.AdamW([{"params": list(m.parameters()), "weight_decay": 5e-4}], lr=1e-3, weight_decay=0.0)
    for q in m.parameters():
        opt.state[q] = {"step": 3, "exp_avg": torch.full_like(q, 0.125), "exp_avg_sq": torch.full_like(q, 0.75)}
    args = _args()
    rutils.save_model(args, (224, 224), 5, m, opt, P.NativeScaler(), ema, 2)
    assert os.path.exists("train_cls/output/checkpoint-5.pth")
    m2, args2 = _perturbed(2, 5), _args()
    ema2 = P.ModelEmaV3(m2, decay=0.9995)
    opt2 = PO.AdamW([{"params": list(m2.parameters()), "weight_decay": 5e-4}], lr=1e-3, weight_decay=0.0)
    rutils.auto_load_model(args2, m2, opt2, P.NativeScaler(), ema2)
    assert _same(m, m2) and _same(ema.module, ema2.module) and args2.start_epoch == 6
    assert all(int(s["step"]) == 3 and torch.equal(s["exp_avg_sq"], torch.full_like(s["exp_avg_sq"], 0.75)) for s in opt2.state.values())
    # and the package's loader reads what the reference wrote, and vice versa
    m3 = _perturbed(2, 6)
    PU.auto_load_model(_args(), m3, PO.AdamW(m3.parameters()), P.NativeScaler(), None, output_dir="train_cls/output")
    assert _same(m, m3)


@pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference is only present in the build container")
def test_reference_val_initialize_model_loads_product_checkpoint(tmp_path):
    """/root/reference/val.py:14-28 unmodified: a checkpoint written by this package's save_model comes back as the pickled
    product model, optionally through the (stand-in) timm ModelEmaV3 with the saved EMA state."""
    pytest.importorskip("torchvision")
    pytest.importorskip("PIL")
    from oracle import shim
    shim.install()
    import sys
    sys.modules["timm.utils"].ModelEmaV3 = P.ModelEmaV3          # the drop-in EMA class, as INTEGRATION.md's patch wires it
    try:
        val = shim._load("_reference_val", f"{REF}/val.py")
    finally:
        from oracle import ema as OE
        sys.modules["timm.utils"].ModelEmaV3 = OE.ModelEmaV3
    m = _perturbed(2, 7)
    ema = P.ModelEmaV3(m, decay=0.9995)
    with torch.no_grad():
        for q in ema.module.parameters():
            q.mul_(0.5)
    opt = PO.AdamW(m.parameters())
    path = PU.save_model(_args(), (224, 224), 2, m, opt, P.NativeScaler(), ema, 2, output_dir=tmp_path)
    net, k = val.initialize_model(str(path), True, "cpu")
    assert k == 2 and isinstance(net, P.ConvNeXt) and not net.training and _same(net, ema.module)
    net, k = val.initialize_model(str(path), False, "cpu")
    assert k == 2 and _same(net, m)
