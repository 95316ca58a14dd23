"""Host-side mirrors of the reference's `utils.py` pieces that sit on the step path or touch this package's objects:

  get_grad_norm_ / clip_grad_norm_        utils.py:456-468 and the `torch.nn.utils.clip_grad_norm_` call of utils.py:440
  NativeScalerWithGradNormCount           utils.py:427-453 (the `loss_scaler` engine.train_one_epoch drives at engine.py:61-68)
  save_model / auto_load_model            utils.py:536-615 (whole-module pickle + state-dict key/shape filter on resume)
  initialize_model                        val.py:14-28 (checkpoint -> eval model, optionally through a fresh ModelEmaV3)

The gradient norm is ONE libcnx launch over a device pointer table (sum of squares per 8192-element chunk, fixed-order final
sum -> deterministic) plus one single-CTA finish kernel that also forms the clip coefficient; clipping is one more launch
that scales every gradient by that device scalar — no host synchronisation anywhere.
"""
from __future__ import annotations

import glob
import os
from pathlib import Path

import torch

from . import _lib as L
from .ema import ModelEmaV3, build_pointer_table, get_state_dict

_NORM_TABLES: dict = {}          # key: tuple of (grad ptr, numel) -> (device table, total chunks, n)


def _grad_table(grads):
    key = tuple((g.data_ptr(), g.numel()) for g in grads)
    hit = _NORM_TABLES.get(key)
    if hit is None:
        if len(_NORM_TABLES) > 8:
            _NORM_TABLES.clear()
        entries, chunk = [], 0
        for g in grads:
            if g.dtype != torch.float32 or not g.is_contiguous():
                raise TypeError("libcnx gradient norm: contiguous fp32 gradients expected (the reference keeps fp32 master grads)")
            entries.append(L.EmaEntry(g.data_ptr(), None, g.numel(), chunk))
            chunk += (g.numel() + L.CNX_EMA_CHUNK - 1) // L.CNX_EMA_CHUNK
        hit = _NORM_TABLES[key] = (build_pointer_table(entries, L.EmaEntry, grads[0].device), chunk, len(entries))
    return hit


def _norm_and_coef(grads, max_norm: float):
    """-> (total_norm, clip_coef) 0-dim fp32 device tensors; clip_coef = min(1, max_norm / (norm + 1e-6)) as torch clips."""
    lib = L.load()
    dev = grads[0].device
    table, chunks, n = _grad_table(grads)
    partial = torch.empty(chunks, dtype=torch.float32, device=dev)
    out = torch.empty(2, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.cnx_grad_sumsq_multi(L.ptr(table), n, chunks, L.ptr(partial), float(max_norm), L.ptr(out), L.stream(dev)),
                "grad_sumsq_multi")
    return out[0], out[1], (table, chunks, n)


def get_grad_norm_(parameters, norm_type: float = 2.0) -> torch.Tensor:
    """utils.py:456-468: the 2-norm of all gradients as a 0-dim device tensor (other norms fall back to the reference's
    torch expression, they are not on the hot path)."""
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads:
        return torch.tensor(0.0)
    if float(norm_type) != 2.0:
        return torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g.detach(), norm_type) for g in grads]), norm_type)
    L.require_cuda(*grads, same_device=False)
    return _norm_and_coef(grads, 0.0)[0]


def clip_grad_norm_(parameters, max_norm: float, norm_type: float = 2.0) -> torch.Tensor:
    """torch.nn.utils.clip_grad_norm_ (utils.py:440) on the device: returns the total norm BEFORE clipping (0-dim tensor) and
    scales every gradient in place by min(1, max_norm / (norm + 1e-6))."""
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    params = [p for p in parameters if p.grad is not None]
    grads = [p.grad for p in params]
    if not grads:
        return torch.tensor(0.0)
    if float(norm_type) != 2.0:
        return torch.nn.utils.clip_grad_norm_(params, max_norm, norm_type)
    L.require_cuda(*grads, same_device=False)
    norm, coef, (table, chunks, n) = _norm_and_coef(grads, float(max_norm))
    lib = L.load()
    with torch.cuda.device(grads[0].device):
        L.check(lib.cnx_scale_multi(L.ptr(table), n, chunks, L.ptr(coef), L.stream(grads[0].device)), "scale_multi")
    torch.autograd.graph.increment_version(grads)
    return norm


class NativeScalerWithGradNormCount:
    """utils.py:427-453 for bf16 autocast: same call signature, same returned gradient norm, `state_dict` /
    `load_state_dict` for the checkpoint's "scaler" entry — without a GradScaler, because bf16 has fp32's exponent range
    (the reference's scaler exists for fp16).  A state dict written by torch's GradScaler loads (and is ignored)."""
    state_dict_key = "amp_scaler"

    def __init__(self):
        self._state = {}

    def __call__(self, loss, optimizer, clip_grad=None, parameters=None, create_graph=False, update_grad=True):
        loss.backward(create_graph=create_graph)
        if not update_grad:
            return None
        if parameters is None:
            raise ValueError("NativeScalerWithGradNormCount: `parameters` is needed for the gradient norm")
        parameters = list(parameters)
        # engine.py:64 passes clip_grad=max_norm whose default is 0: `clip_grad is not None` would then zero every
        # gradient in the reference (coef 0/(norm+1e-6)); train.py:51's default None means "norm only".  0 is treated as None.
        norm = clip_grad_norm_(parameters, clip_grad) if clip_grad else get_grad_norm_(parameters)
        optimizer.step()
        return norm

    def state_dict(self):
        return dict(self._state)

    def load_state_dict(self, state_dict):
        self._state = dict(state_dict)


NativeScaler = NativeScalerWithGradNormCount          # train.py:17 imports it under this name


def save_model(args, input_shape, epoch, model, optimizer, loss_scaler, model_ema, num_classes, output_dir="./train_cls/output"):
    """utils.py:536-553: the checkpoint holds the WHOLE model object (pickled), the optimizer / scaler state dicts and the EMA
    state dict.  Returns the path written."""
    output_dir = Path(output_dir)
    output_dir.mkdir(parents=True, exist_ok=True)
    path = output_dir / ("checkpoint-%s.pth" % str(epoch))
    to_save = {"model": model, "optimizer": optimizer.state_dict(), "epoch": epoch, "scaler": loss_scaler.state_dict(),
               "input_shape": input_shape, "num_classes": num_classes, "args": args}
    if model_ema is not None:
        to_save["model_ema"] = get_state_dict(model_ema)
    torch.save(to_save, path)
    return path


def filter_state_dict(state_dict, model_state_dict):
    """utils.py:584-595: keep the entries whose key exists in the model with the same shape; -> (kept, number skipped)."""
    kept, skipped = {}, 0
    for k, v in state_dict.items():
        if k in model_state_dict and v.shape == model_state_dict[k].shape:
            kept[k] = v
        else:
            print(f"Skipping mismatched key: {k}")
            skipped += 1
    return kept, skipped


def auto_load_model(args, model_without_ddp, optimizer, loss_scaler, model_ema=None, output_dir="./train_cls/output"):
    """utils.py:561-615: resume from `args.resume` (or the newest checkpoint-<n>.pth with `args.auto_resume`)."""
    if getattr(args, "auto_resume", False) and len(getattr(args, "resume", "") or "") == 0:
        latest = -1
        for ckpt in glob.glob(os.path.join(str(output_dir), "checkpoint-*.pth")):
            t = ckpt.split("-")[-1].split(".")[0]
            if t.isdigit():
                latest = max(int(t), latest)
        if latest >= 0:
            args.resume = os.path.join(str(output_dir), "checkpoint-%d.pth" % latest)
    if not getattr(args, "resume", ""):
        return
    checkpoint = torch.load(args.resume, map_location="cpu", weights_only=False)
    state_dict = checkpoint["model"].state_dict()
    kept, missing = filter_state_dict(state_dict, model_without_ddp.state_dict())
    model_without_ddp.load_state_dict(kept, strict=False)
    if getattr(args, "model_ema", False) and model_ema is not None:
        if "model_ema" in checkpoint and missing == 0:
            model_ema.module.load_state_dict(checkpoint["model_ema"])
        else:
            model_ema.set(model_without_ddp)
    if "optimizer" in checkpoint and "epoch" in checkpoint and missing == 0:
        optimizer.load_state_dict(checkpoint["optimizer"])
        if not isinstance(checkpoint["epoch"], str):
            args.start_epoch = checkpoint["epoch"] + 1
        if "scaler" in checkpoint and loss_scaler is not None:
            loss_scaler.load_state_dict(checkpoint["scaler"])


def initialize_model(model_weight_path, model_ema: bool, device):
    """val.py:14-28: -> (eval-ready model, num_classes) from a checkpoint written by save_model."""
    checkpoint = torch.load(model_weight_path, map_location=device, weights_only=False)
    num_classes = checkpoint["num_classes"]
    model = checkpoint["model"]
    if model_ema:
        ema = ModelEmaV3(model, decay=0.999, device=device)
        ema.module.load_state_dict(checkpoint["model_ema"] if "model_ema" in checkpoint else checkpoint["model"].state_dict())
        return ema.module, num_classes
    return model, num_classes
