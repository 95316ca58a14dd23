"""imageclassification_b200 — B200-native (sm_100a) ConvNeXt training-step hot path behind the reference's
Python protocol: `create_model`, `SoftTargetCrossEntropy`, `ModelEmaV3`, `Mixup` are drop-ins for
the timm objects train.py:187-201,256-257 builds and engine.train_one_epoch consumes.  All arithmetic of the hot
path runs in libcnx.so (hand-written CUDA behind the C-ABI of include/cnx.h); there is no CPU fallback."""
from .ema import ModelEmaV3, get_state_dict
from .loss import SoftTargetCrossEntropy
from .mixup import Mixup, mixup_target
from .modules import ConvNeXt, ConvNeXtBlock, LayerNorm, LayerNorm2d, create_model, list_models
from .utils import NativeScaler, NativeScalerWithGradNormCount, clip_grad_norm_, get_grad_norm_

ModelEma = ModelEmaV3

__all__ = ["ConvNeXt", "ConvNeXtBlock", "LayerNorm", "LayerNorm2d", "create_model", "list_models",
           "SoftTargetCrossEntropy", "ModelEmaV3", "ModelEma", "get_state_dict", "Mixup", "mixup_target",
           "NativeScaler", "NativeScalerWithGradNormCount", "clip_grad_norm_", "get_grad_norm_"]
