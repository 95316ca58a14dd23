"""Generate the golden fixtures that pin the oracle against the REFERENCE'S OWN CODE.  Run in the build container only
(needs /root/reference, read-only):   python tests/golden/make_golden.py

  block_C{C}_H{H}.npz   inputs, weights, outputs and all gradients of the reference's in-tree Block
                        (semantic_segmentation/backbone/convnext.py:21-56, imported unmodified through oracle/shim.py)
  ln_channels_first.npz the reference LayerNorm, both data formats (:158-182)
  engine_step.npz       loss / class_acc / per-tensor norms of the parameters and of the EMA after two iterations of the
                        reference's own engine.train_one_epoch (engine.py:10-143, imported unmodified) on BASELINE
                        config 1 (ConvNeXt-T, fp32, batch 8, 2 classes, CPU, seed 88, mixup 0.8, smoothing 0.1, AdamW,
                        EMA 0.9995) — the model/criterion/EMA/mixup objects handed to it are the oracle restatements,
                        because timm itself is not installable here.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import convnext as OC, ema as OE, loss as OL, mixup as OM, shim  # noqa: E402

MAP = {"conv_dw.weight": "dwconv.weight", "conv_dw.bias": "dwconv.bias", "norm.weight": "norm.weight", "norm.bias": "norm.bias",
       "mlp.fc1.weight": "pwconv1.weight", "mlp.fc1.bias": "pwconv1.bias", "mlp.fc2.weight": "pwconv2.weight",
       "mlp.fc2.bias": "pwconv2.bias", "gamma": "gamma"}


def block_golden(ref, C, H, N, seed, gamma_init):
    torch.manual_seed(seed)
    blk = ref.Block(C, drop_path=0.0, layer_scale_init_value=gamma_init)
    with torch.no_grad():
        for n, p in blk.named_parameters():
            if n.endswith("bias"):
                p.normal_(0, 0.05)
            elif n == "norm.weight":
                p.add_(0.1 * torch.randn_like(p))
            elif n == "gamma":
                p.mul_(1 + 0.2 * torch.randn_like(p))
            else:
                p.normal_(0, 0.05)
    x = torch.randn(N, C, H, H, requires_grad=True)
    dout = torch.randn(N, C, H, H)
    y = blk(x)
    y.backward(dout)
    out = {"x": x.detach().numpy(), "dout": dout.numpy(), "y": y.detach().numpy(), "dx": x.grad.numpy()}
    sd = dict(blk.named_parameters())
    for ours, theirs in MAP.items():
        out["p." + ours] = sd[theirs].detach().numpy()
        out["g." + ours] = sd[theirs].grad.numpy()
    np.savez_compressed(os.path.join(HERE, f"block_C{C}_H{H}.npz"), **out)
    print("block", C, H, float(y.abs().mean()))


def ln_golden(ref):
    torch.manual_seed(3)
    C = 96
    w, b = 1 + 0.1 * torch.randn(C), 0.1 * torch.randn(C)
    x = torch.randn(2, 5, 7, C)
    cl = ref.LayerNorm(C, eps=1e-6)
    cf = ref.LayerNorm(C, eps=1e-6, data_format="channels_first")
    with torch.no_grad():
        for m in (cl, cf):
            m.weight.copy_(w)
            m.bias.copy_(b)
        y_cl = cl(x)
        y_cf = cf(x.permute(0, 3, 1, 2))
    np.savez_compressed(os.path.join(HERE, "ln_channels_first.npz"), x=x.numpy(), w=w.numpy(), b=b.numpy(), y_cl=y_cl.numpy(),
                        y_cf=y_cf.numpy())
    print("ln", float((y_cl - y_cf.permute(0, 2, 3, 1)).abs().max()))


def engine_golden():
    eng = shim.import_reference_engine()
    torch.cuda.synchronize = lambda *a, **k: None        # engine.py:79 raises on a driver-less host
    out = {}
    for tag, img, gamma_init, dpr in (("a", 224, 1e-6, 0.05), ("b", 64, 1.0, 0.0)):
        torch.manual_seed(88)
        np.random.seed(88)
        model = OC.create_model("convnext_tiny", num_classes=2, drop_path_rate=dpr, ls_init_value=gamma_init)
        ema = OE.ModelEmaV3(model, decay=0.9995, device=torch.device("cpu"))
        opt = torch.optim.AdamW([{"params": list(model.parameters()), "weight_decay": 5e-4}], lr=1e-3, weight_decay=0.0)
        mix = OM.Mixup(mixup_alpha=0.8, cutmix_alpha=0.0, label_smoothing=0.1, num_classes=2)
        data = [(torch.randn(8, 3, img, img), torch.randint(0, 2, (8,))) for _ in range(2)]
        stats = eng.train_one_epoch(model, OL.SoftTargetCrossEntropy(), data, opt, torch.device("cpu"), 0, None, None, ema, mix,
                                    start_steps=0, num_training_steps_per_epoch=2, update_freq=1, use_amp=False, num_classes=2)
        out[f"{tag}.loss"] = np.float64(stats["loss"])
        out[f"{tag}.class_acc"] = np.float64(stats["class_acc"])
        out[f"{tag}.param_norms"] = np.array([p.detach().double().norm().item() for p in model.parameters()])
        out[f"{tag}.param_sums"] = np.array([p.detach().double().sum().item() for p in model.parameters()])
        out[f"{tag}.ema_norms"] = np.array([p.detach().double().norm().item() for p in ema.module.parameters()])
        out[f"{tag}.ema_sums"] = np.array([p.detach().double().sum().item() for p in ema.module.parameters()])
        print("engine", tag, stats)
    np.savez_compressed(os.path.join(HERE, "engine_step.npz"), **out)


def evaluate_golden():
    """evaluate.npz: the dict the reference's own engine.evaluate (engine.py:145-225, imported unmodified) returns for a seeded
    ConvNeXt-T (3 classes, uneven batch sizes so the batch-mean loss and the sample-weighted acc1 differ), CPU fp32."""
    eng = shim.import_reference_engine()
    torch.manual_seed(5)
    model = OC.create_model("convnext_tiny", num_classes=3, ls_init_value=1.0)
    with torch.no_grad():
        model.head.fc.weight.normal_(0, 0.5)          # spread the logits so that all three classes get predicted
    data = [(torch.randn(b, 3, 64, 64), torch.randint(0, 3, (b,))) for b in (8, 8, 5)]
    stats = eng.evaluate(data, model, torch.device("cpu"), 3, use_amp=False)
    np.savez_compressed(os.path.join(HERE, "evaluate.npz"), keys=np.array(sorted(stats)),
                        values=np.array([float(stats[k]) for k in sorted(stats)]))
    print("evaluate", stats)


if __name__ == "__main__":
    torch.set_num_threads(8)
    ref = shim.import_reference_backbone()
    block_golden(ref, 32, 9, 2, 1, 1.0)
    block_golden(ref, 96, 7, 2, 2, 1.0)
    block_golden(ref, 64, 12, 1, 3, 1e-6)
    ln_golden(ref)
    engine_golden()
    evaluate_golden()
