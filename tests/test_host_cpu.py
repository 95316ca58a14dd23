"""CPU-side checks of the product: the C-ABI library loads and exports every symbol include/cnx.h declares, argument
validation works without a device, the host mirror keeps the reference's interface (names, state-dict keys, RNG call
sequence, picklability) and fails loudly — never falls back — when handed CPU tensors."""
import copy
import os
import pickle
import re

import numpy as np
import pytest
import torch

import imageclassification_b200 as P
from imageclassification_b200 import _lib as L
from oracle import convnext as OC, mixup as OM

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "cnx.h")).read()
    declared = set(re.findall(r"CNX_API\s+[\w\s\*]+?\b(cnx_\w+)\s*\(", hdr))
    assert len(declared) >= 25
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)
    lib = L.load()
    for name in declared:
        assert hasattr(lib.cdll, name), name
    assert lib.cnx_version() == 100
    assert set(L.KERNELS_PER_CALL) <= declared


def test_argument_validation_needs_no_device():
    lib = L.load()
    assert lib.cnx_gemm_plain(None, None, None, None, 0, 8, 8, 8, 0, 0, None) == -1
    assert b"null pointer" in lib.cnx_last_error_string()
    assert lib.cnx_ema_lerp_multi(None, 0, 0, 0.5, None) == -1
    assert lib.cnx_mixup_target(None, 4, 4, 0.5, 0.1, None, None) == -1
    assert lib.cnx_gemm_wgrad_workspace_bytes(0, 8, 8, 1, 0) == 0
    # entry points added for the image mixing, the fp32-accurate (x3) forward and the head: shape / alignment errors are
    # reported (negative CNX_E_*) before anything is launched
    fake = 4096                                             # a non-null, 16-byte aligned "pointer" that is never dereferenced
    assert lib.cnx_mixup_batch(None, None, 4, 3, 8, 8, 0.5, 0, 0, 0, 0, 0, None) == -1
    assert lib.cnx_mixup_batch(fake, None, 4, 3, 8, 8, 0.5, 1, 0, 9, 0, 1, None) == -1 and b"cutmix box" in lib.cnx_last_error_string()
    assert lib.cnx_mixup_batch(fake + 4, None, 4, 3, 8, 8, 0.5, 0, 0, 0, 0, 0, None) == -1 and b"aligned" in lib.cnx_last_error_string()
    assert lib.cnx_split3(fake, 16, 12, fake, 3, None) == -2 and b"multiple of 8" in lib.cnx_last_error_string()
    assert lib.cnx_split3(fake, 16, 16, fake, 4, None) == -1
    assert lib.cnx_gemm_bias_gelu_fwd_x3(fake, fake, fake, 128, 48, 72, fake, 3, None) == -2 and b"multiple of 32" in lib.cnx_last_error_string()
    assert lib.cnx_gemm_bias_gelu_fwd_x3(fake, fake, fake, 128, 64, 70, fake, 3, None) == -1
    assert lib.cnx_gemm_bias_gelu_fwd_x3(fake, fake, fake, 128, 64, 72, fake, 2, None) == -2 and b"a_segments" in lib.cnx_last_error_string()
    assert lib.cnx_gemm_bias_scale_residual_fwd(fake, fake, None, None, None, 49, fake, fake, 0, 128, 96, 300, 1, L.CNX_GEMM_A_SPLIT2, None) == -2
    assert lib.cnx_dwconv7_ln_fwd_x3(fake, fake, fake, fake, fake, 1e-6, 1, 7, 7, 40, fake, fake, fake, fake, 3, None) == -2
    assert lib.cnx_dwconv7_ln_fwd_x3(fake, fake, fake, fake, fake, 1e-6, 1, 7, 7, 64, fake, fake, fake, fake, 1, None) == -1
    assert lib.cnx_avgpool_nhwc_fwd(fake, 0, 2, 49, 6, fake, None) == -2 and lib.cnx_avgpool_nhwc_bwd(None, 2, 49, 8, fake, 0, None) == -1
    assert lib.cnx_weight_prep(fake, 8, 8, None, 3, fake, 0, None) == -1 and b"mode 3" in lib.cnx_last_error_string()


def test_no_cpu_fallback():
    blk = P.ConvNeXtBlock(32)
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        blk(torch.randn(1, 32, 8, 8))
    with pytest.raises(RuntimeError):
        P.SoftTargetCrossEntropy()(torch.zeros(2, 3), torch.zeros(2, 3))
    with pytest.raises(RuntimeError):
        P.mixup_target(torch.zeros(2, dtype=torch.long), 3, 0.5, 0.1)
    lin = torch.nn.Linear(4, 4)
    ema = P.ModelEmaV3(lin, decay=0.9)
    with pytest.raises(RuntimeError):
        ema.update(lin)
    from imageclassification_b200 import engine
    with pytest.raises(RuntimeError):
        engine.train_one_epoch(lin, None, [], torch.optim.SGD(lin.parameters(), lr=0.1), "cpu", 0)


def test_model_interface_matches_timm_names():
    m = P.create_model("convnext_tiny", pretrained=False, num_classes=2, drop_path_rate=0.05)
    o = OC.create_model("convnext_tiny", num_classes=2, drop_path_rate=0.05)
    sd, so = m.state_dict(), o.state_dict()
    assert list(sd) == list(so)
    assert all(sd[k].shape == so[k].shape for k in sd)
    m.load_state_dict(so)                                              # oracle/timm-format checkpoints load
    assert sum(p.numel() for p in m.parameters()) == 27_821_666
    rates = [b.drop_path.drop_prob if hasattr(b.drop_path, "drop_prob") else 0.0 for s in m.stages for b in s.blocks]
    assert rates[0] == 0.0 and abs(rates[-1] - 0.05) < 1e-7 and len(rates) == 18     # convnext.py:92 linear schedule
    m2 = pickle.loads(pickle.dumps(m))                                 # utils.py:542 pickles whole modules
    assert list(m2.state_dict()) == list(sd)
    e = P.ModelEmaV3(m, decay=0.9995)
    assert not e.module.training and list(P.get_state_dict(e)) == list(sd)
    pickle.loads(pickle.dumps(e))
    copy.deepcopy(m)
    with pytest.raises(ValueError):
        P.create_model("resnet50")
    with pytest.raises(RuntimeError):
        P.create_model("convnext_tiny", pretrained=True)
    for name in ("convnext_base", "convnext_large"):
        assert name in P.list_models()


def test_mixup_host_rng_sequence_matches_oracle():
    for kw in [dict(mixup_alpha=0.8), dict(mixup_alpha=0.8, cutmix_alpha=1.0), dict(mixup_alpha=0.0, cutmix_alpha=1.0),
               dict(mixup_alpha=0.0, cutmix_minmax=(0.2, 0.8)), dict(mixup_alpha=0.8, cutmix_alpha=1.0, prob=0.3)]:
        a, b = P.Mixup(num_classes=10, **kw), OM.Mixup(num_classes=10, **kw)
        for seed in range(20):
            np.random.seed(seed)
            ra = a._params_per_batch()
            sa = np.random.get_state()[1][:4].tolist()
            np.random.seed(seed)
            rb = b._params_per_batch()
            assert ra == rb and sa == np.random.get_state()[1][:4].tolist()
    from imageclassification_b200 import mixup as PM
    for seed in range(10):
        np.random.seed(seed)
        r1 = PM.cutmix_bbox_and_lam((8, 3, 224, 224), 0.3)
        np.random.seed(seed)
        r2 = OM.cutmix_bbox_and_lam((8, 3, 224, 224), 0.3)
        assert r1 == r2
    with pytest.raises(NotImplementedError):
        P.Mixup(mode="elem")


def test_torch_library_registration():
    """BASELINE north_star: "the Python host code calls the kernels as torch custom ops": every entry point of the hot path is
    registered with torch.library under torch.ops.cnx and is what the nn.Module mirror calls."""
    import inspect
    from imageclassification_b200 import loss, mixup, modules, ops
    for name in ("block_forward", "layer_norm_cl", "stem_forward", "downsample_forward", "head_forward",
                 "soft_target_cross_entropy", "mixup_target", "mixup_batch"):
        op = getattr(torch.ops.cnx, name)
        assert "cnx::" + name in str(op.default._schema)
        assert name in ops._SCHEMAS
    assert "Tensor? dp" in str(torch.ops.cnx.block_forward.default._schema)
    assert "Tensor(a!) x" in str(torch.ops.cnx.mixup_batch.default._schema)
    for mod in (modules, loss, mixup):
        assert "torch.ops.cnx." in inspect.getsource(mod)
    with pytest.raises(RuntimeError, match="CUDA"):                       # dispatched, then refused: no CPU fallback behind the op
        torch.ops.cnx.mixup_target(torch.zeros(2, dtype=torch.long), 3, 0.5, 0.1)
