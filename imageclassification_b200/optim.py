"""AdamW — `torch.optim.AdamW` as the reference builds it (optim_factory.py:74-75: lr, weight_decay, default betas/eps,
one parameter group holding every parameter, optim_factory.py:23-47), with the whole step in ONE libcnx launch over a
device pointer table, optionally fused with the ModelEmaV3 update that engine.py:73-77 runs right after it
(36*P bytes in one pass instead of 28*P + 12*P and hundreds of multi-tensor-apply chunks).

Arithmetic follows torch/optim/adamw.py single-tensor form, in fp32:
    p *= 1 - lr*wd;  m = lerp(m, g, 1-b1);  v = b2*v + (1-b2)*g*g
    p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps);   [ema = lerp(ema, p, 1-decay)]
"""
from __future__ import annotations

import ctypes
import math

import numpy as np
import torch

from . import _lib as L


class _TableUploader:
    """Pointer tables go host->device through a small ring of pinned buffers with non-blocking copies, so a table
    rebuild (gradient tensors re-allocated after zero_grad(set_to_none=True)) never stalls the launching thread."""

    def __init__(self, slots: int = 4):
        self.slots = [None] * slots
        self.events = [None] * slots
        self.i = 0

    def upload(self, raw: bytes, device) -> torch.Tensor:
        n = len(raw)
        s = self.i
        self.i = (self.i + 1) % len(self.slots)
        if self.slots[s] is None or self.slots[s].numel() < n:
            self.slots[s] = torch.empty(max(n, 1 << 16), dtype=torch.uint8).pin_memory()
            self.events[s] = None
        if self.events[s] is not None:
            self.events[s].synchronize()
        self.slots[s][:n].copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
        dev = torch.empty(n, dtype=torch.uint8, device=device)
        dev.copy_(self.slots[s][:n], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.events[s] = ev
        return dev


class AdamW(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 amsgrad: bool = False):
        if amsgrad:
            raise NotImplementedError("amsgrad is not implemented (the reference never enables it)")
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid AdamW hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._ema = None
        self._ema_map = {}
        self._uploader = _TableUploader()
        self._tables = {}

    def fuse_ema(self, model_ema, model) -> None:
        """Fold `model_ema.update(model)` into step(): the EMA tensor of every optimised parameter is updated in the
        same kernel; the next `model_ema.update()` call only handles what the optimiser does not own."""
        base = model.module if hasattr(model, "module") and not hasattr(model, "stem") else model
        ema_params = dict(model_ema.module.named_parameters())
        self._ema_map = {p: ema_params[n] for n, p in base.named_parameters() if n in ema_params}
        self._ema = model_ema
        self._ema_ptrs = frozenset(e.data_ptr() for e in self._ema_map.values())
        self._tables = {}
        object.__setattr__(model_ema, "_fused_optimizer", self)

    # the device pointer tables are caches over (param, grad, exp_avg, exp_avg_sq, ema) addresses: anything that can replace
    # one of those tensors drops them
    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._tables = {}

    def add_param_group(self, param_group):
        super().add_param_group(param_group)
        self._tables = {}

    def __setstate__(self, state):
        super().__setstate__(state)
        self._tables = {}
        if not hasattr(self, "_uploader"):
            self._uploader = _TableUploader()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = L.load()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            L.require_cuda(*ps, same_device=False)             # launched under torch.cuda.device(...) below
            by_step = {}                                           # per-parameter step counts, as torch keeps them
            for p in ps:
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                elif isinstance(st["step"], torch.Tensor):          # a state dict written by torch.optim.AdamW
                    st["step"] = int(st["step"].item())
                st["step"] += 1
                by_step.setdefault(st["step"], []).append(p)
            b1, b2 = group["betas"]
            ema_w = (1.0 - self._ema.get_decay()) if self._ema is not None else 0.0
            for si, (t, sub) in enumerate(sorted(by_step.items())):
                # one launch per distinct step count: normally one (every parameter receives a gradient every step)
                key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr(), self.state[p]["exp_avg_sq"].data_ptr(),
                             self._ema_map[p].data_ptr() if p in self._ema_map else 0) for p in sub)
                cached = self._tables.get((gi, si))
                if cached is None or cached[0] != key:
                    entries, chunk = [], 0
                    for p in sub:
                        st = self.state[p]
                        for q in (p, p.grad, st["exp_avg"], st["exp_avg_sq"]):
                            if q.dtype != torch.float32 or not q.is_contiguous() or q.device != p.device:
                                raise TypeError("AdamW: libcnx updates contiguous fp32 parameters, gradients and moments on one device")
                        e = self._ema_map.get(p)
                        entries.append(L.AdamWEntry(p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(),
                                                    st["exp_avg_sq"].data_ptr(), e.data_ptr() if e is not None else None,
                                                    p.numel(), chunk))
                        chunk += (p.numel() + L.CNX_EMA_CHUNK - 1) // L.CNX_EMA_CHUNK
                    arr = (L.AdamWEntry * len(entries))(*entries)
                    table = self._uploader.upload(bytes(arr), sub[0].device)
                    cached = (key, table, chunk, len(entries))
                    self._tables[(gi, si)] = cached
                _, table, chunks, n = cached
                bc1 = 1.0 - b1 ** t
                bc2_sqrt = math.sqrt(1.0 - b2 ** t)
                with torch.cuda.device(sub[0].device):
                    L.check(lib.cnx_adamw_ema_multi(L.ptr(table), n, chunks, float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                                    float(group["weight_decay"]), bc1, bc2_sqrt, ctypes.c_float(ema_w),
                                                    L.stream(sub[0].device)), "adamw_ema_multi")
            # the kernel wrote parameters and EMA tensors through raw pointers: tell autograd / the derived-weight caches
            torch.autograd.graph.increment_version(ps)
            emas = [self._ema_map[p] for p in ps if p in self._ema_map]
            if emas:
                torch.autograd.graph.increment_version(emas)
        if self._ema is not None:
            self._ema._fused_done = True
        return loss
