#!/usr/bin/env python
"""Context number: the reference's training step built from stock torch.nn modules (the oracle objects: ATen /
cuDNN / cuBLAS kernels, torch.optim.AdamW, foreach-lerp EMA) on the SAME B200, same workload as bench.py's headline
(ConvNeXt-T 224^2, batch 256, 1000 classes, mixup 0.8, smoothing 0.1, SoftTargetCE, EMA 0.9995, accuracy forward outside autocast
as engine.py:89-97 has it).  `--amp` = torch.autocast(bf16), default fp32 (the reference's --use_amp false).
Prints one JSON line.   usage: python profiles/torch_gpu_step.py [--amp] [--steps 10] [--batch 256]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import convnext as OC, ema as OE, engine as OEng, loss as OL, mixup as OM  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--amp", action="store_true")
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--model", default="convnext_tiny")
ap.add_argument("--channels-last", action="store_true")
ap.add_argument("--fast-loop", action="store_true",
                help="drive the stock modules with THIS package's step loop (device-side class counts, no 3*K .item() syncs, loss "
                     "read back after backward, side-stream prefetch): the honest stock-kernel GPU baseline")
a = ap.parse_args()
dev = torch.device("cuda")
torch.manual_seed(88)
np.random.seed(88)
model = OC.create_model(a.model, num_classes=1000, drop_path_rate=0.05).to(dev)
if a.channels_last:
    model = model.to(memory_format=torch.channels_last)
ema = OE.ModelEmaV3(model, decay=0.9995, device=dev)
opt = torch.optim.AdamW([{"params": list(model.parameters()), "weight_decay": 5e-4}], lr=1e-3, weight_decay=0.0)
mix = OM.Mixup(mixup_alpha=0.8, label_smoothing=0.1, num_classes=1000)
crit = OL.SoftTargetCrossEntropy()
g = torch.Generator().manual_seed(1)
data = [(torch.randn(a.batch, 3, 224, 224, generator=g).to(dev), torch.randint(0, 1000, (a.batch,), generator=g).to(dev)) for _ in range(2)]


def epoch(n):
    if a.fast_loop:
        from imageclassification_b200 import engine as PE
        PE.train_one_epoch(model, crit, [data[i % 2] for i in range(n)], opt, dev, 0, None, 0, ema, mix, use_amp=a.amp, num_classes=1000,
                           verbose=False)
    else:
        OEng.train_one_epoch(model, crit, [data[i % 2] for i in range(n)], opt, dev, 0, None, 0, ema, mix, use_amp=a.amp, num_classes=1000)


epoch(3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
epoch(a.steps)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
print(json.dumps({"impl": "stock torch.nn modules on the GPU (oracle objects + oracle step loop)", "model": a.model, "batch": a.batch,
                  "amp_bf16": a.amp, "channels_last": a.channels_last, "ms_per_step": round(ms, 2),
                  "images_per_s": round(a.batch * 1e3 / ms, 1),
                  "step_loop": "imageclassification_b200.engine (no per-class .item() loops)" if a.fast_loop else "oracle/engine.py",
                  "note": "stock ATen / cuDNN / cuBLAS kernels; " + ("device-side class counts, one host read-back of the loss per step"
                          if a.fast_loop else "the oracle step loop keeps engine.py's 3*num_classes .item() syncs per step (3000 at K=1000)")}))
