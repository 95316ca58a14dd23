"""ModelEmaV3 — drop-in for timm.utils.ModelEmaV3 as used by the reference: ctor `ModelEmaV3(model, decay=0.9995,
device=device)` (train.py:201; val.py:19), `.update(model)` after every optimizer step where `model` may be the DDP
wrapper (engine.py:68,77), `.module` for evaluation and checkpointing (train.py:276,367; utils.py:551,601), `.set(model)`
(utils.py:603).  The update of every floating state-dict tensor is ONE libcnx launch over a device-resident pointer
table, bit-exact with ATen lerp on both of its branches: w = fp32(1-decay);
ema <- fmaf(w, p - ema, ema) for |w| < 0.5, p - (p - ema)*(1 - w) otherwise (e.g. w = 1 while get_decay(step) is 0)."""
from __future__ import annotations

import ctypes
from copy import deepcopy

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L


def build_pointer_table(entries, entry_type, device):
    """ctypes array of table entries -> device uint8 tensor (kept alive by the caller)."""
    arr = (entry_type * len(entries))(*entries)
    host = torch.from_numpy(np.frombuffer(bytes(arr), dtype=np.uint8).copy())
    return host.to(device)


class ModelEmaV3(nn.Module):
    def __init__(self, model, decay: float = 0.9999, min_decay: float = 0.0, update_after_step: int = 0,
                 use_warmup: bool = False, warmup_gamma: float = 1.0, warmup_power: float = 2 / 3, device=None,
                 foreach: bool = True, exclude_buffers: bool = False):
        super().__init__()
        self.module = deepcopy(model)
        self.module.eval()
        self.decay = decay
        self.min_decay = min_decay
        self.update_after_step = update_after_step
        self.use_warmup = use_warmup
        self.warmup_gamma = warmup_gamma
        self.warmup_power = warmup_power
        self.foreach = foreach
        self.device = device
        self.exclude_buffers = exclude_buffers
        if self.device is not None and torch.device(device) != next(model.parameters()).device:
            self.module.to(device=device)
        self._table_key = None
        self._table = None
        self._chunks = 0

    def get_decay(self, step=None) -> float:
        if step is None:
            return self.decay
        step = max(0, step - self.update_after_step - 1)
        if step <= 0:
            return 0.0
        if self.use_warmup:
            decay = 1 - (1 + step / self.warmup_gamma) ** -self.warmup_power
            return max(min(decay, self.decay), self.min_decay)
        return self.decay

    def _pairs(self, model):
        if self.exclude_buffers:
            ema = [p for _, p in self.module.named_parameters()]
            mod = [p for _, p in model.named_parameters()]
            bufs = list(zip(self.module.buffers(), model.buffers()))
        else:
            ema = list(self.module.state_dict().values())
            mod = list(model.state_dict().values())
            bufs = []
        if len(ema) != len(mod):
            raise RuntimeError(f"ModelEmaV3.update: EMA has {len(ema)} tensors, model has {len(mod)}")
        return ema, mod, bufs

    @torch.no_grad()
    def update(self, model, step=None):
        decay = self.get_decay(step)
        ema, mod, bufs = self._pairs(model)
        fused = ()
        if getattr(self, "_fused_done", False):
            # imageclassification_b200.optim.AdamW.fuse_ema: the optimizer kernel already moved these EMA tensors
            self._fused_done = False
            fused = self._fused_optimizer._ema_ptrs
        fl_e, fl_m = [], []
        for e, m in zip(ema, mod):
            if fused and e.data_ptr() in fused:
                continue
            if e.is_floating_point():
                fl_e.append(e)
                fl_m.append(m)
            else:
                e.copy_(m)
        for e, m in bufs:
            e.copy_(m)
        if not fl_e:
            return
        L.require_cuda(*fl_e, same_device=False)        # launched under torch.cuda.device(...) below
        key = tuple((e.data_ptr(), m.data_ptr(), e.numel()) for e, m in zip(fl_e, fl_m))
        if key != self._table_key:
            entries, chunk = [], 0
            for e, m in zip(fl_e, fl_m):
                if e.dtype != torch.float32 or m.dtype != torch.float32 or not (e.is_contiguous() and m.is_contiguous()):
                    raise TypeError("ModelEmaV3: libcnx updates contiguous fp32 tensors (the reference keeps fp32 master "
                                    f"weights under autocast); got {e.dtype}/{m.dtype}")
                if m.device != e.device:
                    raise RuntimeError("ModelEmaV3: model and EMA must live on the same CUDA device")
                entries.append(L.EmaEntry(e.data_ptr(), m.data_ptr(), e.numel(), chunk))
                chunk += (e.numel() + L.CNX_EMA_CHUNK - 1) // L.CNX_EMA_CHUNK
            self._table = build_pointer_table(entries, L.EmaEntry, fl_e[0].device)
            self._table_key, self._chunks = key, chunk
        lib = L.load()
        with torch.cuda.device(fl_e[0].device):
            L.check(lib.cnx_ema_lerp_multi(L.ptr(self._table), len(key), self._chunks, ctypes.c_float(1.0 - decay),
                                           L.stream(fl_e[0].device)), "ema_lerp_multi")
        # written through raw pointers: bump the version counters so that layouts derived from the EMA weights
        # (ops._derived / ops._PrepRegistry key on (data_ptr, _version)) are rebuilt before the next forward of `.module`
        torch.autograd.graph.increment_version(fl_e)

    @torch.no_grad()
    def set(self, model):
        for e, m in zip(self.module.state_dict().values(), model.state_dict().values()):
            e.copy_(m.to(device=e.device))

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def __getstate__(self):
        # the device pointer table is a cache, never part of a checkpoint (utils.py:542 pickles whole modules)
        st = super().__getstate__() if hasattr(super(), "__getstate__") else self.__dict__.copy()
        st = dict(st)
        st["_table_key"], st["_table"], st["_chunks"] = None, None, 0
        st.pop("_fused_optimizer", None)
        st.pop("_fused_done", None)
        return st


def get_state_dict(model, unwrap_fn=None):
    """timm.utils.get_state_dict as called at utils.py:551."""
    m = model
    while hasattr(m, "module"):
        m = m.module
    return m.state_dict()
