"""Stand-ins for the third-party names the reference imports, so its OWN files can be imported in the build
container (TEST ORACLE tooling; used only by tests/golden/make_golden.py and tests that skip when
/root/reference is absent).

  engine.py:4-5      timm.data.Mixup ; timm.utils.accuracy, ModelEmaV3
  val.py:9-10        timm.utils.ModelEmaV3 ; timm.data.constants.IMAGENET_DEFAULT_{MEAN,STD}
  utils.py:7,15      timm.utils.get_state_dict ; tensorboardX.SummaryWriter
  semantic_segmentation/backbone/convnext.py:14-18
                     timm.models.layers.{trunc_normal_, DropPath} ; mmcv_custom.load_checkpoint ;
                     mmseg.utils.get_root_logger ; mmseg.models.builder.BACKBONES
Every timm symbol is backed by the oracle restatement in this package — nothing here is product code."""
from __future__ import annotations

import importlib.util
import sys
import types

import torch

from . import convnext as _cn, ema as _ema, loss as _loss, mixup as _mix

REFERENCE_ROOT = "/root/reference"


def _accuracy(output, target, topk=(1,)):
    maxk = min(max(topk), output.size()[1])
    batch_size = target.size(0)
    _, pred = output.topk(maxk, 1, True, True)
    pred = pred.t()
    correct = pred.eq(target.reshape(1, -1).expand_as(pred))
    return [correct[:min(k, maxk)].reshape(-1).float().sum(0) * 100.0 / batch_size for k in topk]


class _Registry:
    def register_module(self, *a, **k):
        return lambda cls: cls


def install():
    """Insert the stand-in modules into sys.modules (idempotent)."""
    def mod(name, **attrs):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    mod("timm")
    mod("timm.data", Mixup=_mix.Mixup)
    mod("timm.data.mixup", Mixup=_mix.Mixup, mixup_target=_mix.mixup_target)
    # val.py:10 — public ImageNet normalisation constants (timm/data/constants.py)
    mod("timm.data.constants", IMAGENET_DEFAULT_MEAN=(0.485, 0.456, 0.406), IMAGENET_DEFAULT_STD=(0.229, 0.224, 0.225))
    mod("timm.utils", accuracy=_accuracy, ModelEmaV3=_ema.ModelEmaV3, get_state_dict=_ema.get_state_dict)
    mod("timm.loss", SoftTargetCrossEntropy=_loss.SoftTargetCrossEntropy)
    mod("timm.models", create_model=_cn.create_model)
    mod("timm.models.layers", trunc_normal_=torch.nn.init.trunc_normal_, DropPath=_cn.DropPath)
    mod("tensorboardX", SummaryWriter=type("SummaryWriter", (), {"__init__": lambda self, *a, **k: None}))
    mod("mmcv_custom", load_checkpoint=lambda *a, **k: None)
    mod("mmseg")
    mod("mmseg.utils", get_root_logger=lambda *a, **k: None)
    mod("mmseg.models")
    mod("mmseg.models.builder", BACKBONES=_Registry())


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def import_reference_engine(root: str = REFERENCE_ROOT):
    """Import the reference's unmodified utils.py and engine.py (read-only) on top of the stand-ins."""
    install()
    if "utils" not in sys.modules or not getattr(sys.modules["utils"], "__file__", "").startswith(root):
        _load("utils", f"{root}/utils.py")
    return _load("_reference_engine", f"{root}/engine.py")


def import_reference_backbone(root: str = REFERENCE_ROOT):
    """Import the reference's in-tree ConvNeXt Block / LayerNorm (semantic_segmentation/backbone/convnext.py)."""
    install()
    return _load("_reference_convnext", f"{root}/semantic_segmentation/backbone/convnext.py")
