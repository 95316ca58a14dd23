"""timm.utils.ModelEmaV3 restated (TEST ORACLE).  Constructed at train.py:201 as ModelEmaV3(model, decay=0.9995,
device=device); .update(model) at engine.py:68,77; .module used at train.py:276,367; .set at utils.py:603.
Published algorithm (timm/utils/model_ema.py): ema <- lerp(ema, param, 1 - decay) over every floating state-dict
value (torch._foreach_lerp_), integer values copied."""
from copy import deepcopy

import torch
import torch.nn as nn


class ModelEmaV3(nn.Module):
    def __init__(self, model, decay=0.9999, min_decay=0.0, update_after_step=0, use_warmup=False, warmup_gamma=1.0,
                 warmup_power=2 / 3, device=None, foreach=True, exclude_buffers=False):
        super().__init__()
        self.module = deepcopy(model)
        self.module.eval()
        self.decay, self.min_decay, self.update_after_step = decay, min_decay, update_after_step
        self.use_warmup, self.warmup_gamma, self.warmup_power = use_warmup, warmup_gamma, warmup_power
        self.foreach, self.device, self.exclude_buffers = foreach, device, exclude_buffers
        if self.device is not None and device != next(model.parameters()).device:
            self.foreach = False
            self.module.to(device=device)

    def get_decay(self, step=None):
        if step is None:
            return self.decay
        step = max(0, step - self.update_after_step - 1)
        if step <= 0:
            return 0.0
        if self.use_warmup:
            decay = 1 - (1 + step / self.warmup_gamma) ** -self.warmup_power
            return max(min(decay, self.decay), self.min_decay)
        return self.decay

    @torch.no_grad()
    def update(self, model, step=None):
        decay = self.get_decay(step)
        ema_f, mod_f = [], []
        for e, m in zip(self.module.state_dict().values(), model.state_dict().values()):
            if e.is_floating_point():
                ema_f.append(e)
                mod_f.append(m.to(device=e.device))
            else:
                e.copy_(m)
        if self.foreach:
            torch._foreach_lerp_(ema_f, mod_f, weight=1.0 - decay)
        else:
            for e, m in zip(ema_f, mod_f):
                e.lerp_(m, weight=1.0 - decay)

    @torch.no_grad()
    def set(self, model):
        for e, m in zip(self.module.state_dict().values(), model.state_dict().values()):
            e.copy_(m.to(device=e.device))

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)


def get_state_dict(model, unwrap_fn=None):
    """timm.utils.get_state_dict as used at utils.py:551: unwrap .module (EMA / DDP) then state_dict()."""
    m = model
    while hasattr(m, "module"):
        m = m.module
    return m.state_dict()
