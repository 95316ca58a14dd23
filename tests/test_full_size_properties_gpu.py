"""Parity at BASELINE.json's FULL sizes (config 2: batch 256, ConvNeXt-T stage shapes, 1000 classes, 28.6 M parameters), where
running the oracle would take minutes: size-independent properties the arithmetic must have — exact homogeneity of the GEMMs
under power-of-two scaling, the adjoint (dot-product) identities that tie every backward kernel to its forward, LayerNorm
moments, probability-simplex invariants of the loss and label mixing, symmetry of the image mixing, fixed points of the EMA.
Everything goes through the C-ABI / the public modules on the device."""
import pytest
import torch

import imageclassification_b200 as P
from cabi import (dwconv7_dgrad, dwconv7_ln_fwd, dwconv7_wgrad, gemm_bias_gelu, gemm_plain, gemm_scale_res, gemm_wgrad, ln_bwd,
                  max_rel)
from imageclassification_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"
N = 256                                    # BASELINE config 2 batch
STAGES = [(96, 56), (384, 14)]             # the HBM-bound and the tensor-bound stage of ConvNeXt-T at 224^2


def _bits(a, b):
    return torch.equal(a.contiguous().view(torch.int16 if a.element_size() == 2 else torch.int32),
                       b.contiguous().view(torch.int16 if b.element_size() == 2 else torch.int32))


@pytest.mark.parametrize("C,H", STAGES)
def test_gemms_are_exactly_homogeneous(C, H):
    """Scaling an operand by 2 is exact in floating point, so every output must scale bit-exactly — any dropped, duplicated or
    mis-addressed tile at M = 802 816 / 50 176 rows breaks it."""
    M = N * H * H
    g = torch.Generator(device=DEV).manual_seed(C)
    A = torch.randn(M, C, device=DEV, generator=g).to(torch.bfloat16)
    W = (torch.randn(4 * C, C, device=DEV, generator=g) / C ** 0.5).to(torch.bfloat16)
    y1 = gemm_plain(A, W, None, torch.bfloat16)
    y2 = gemm_plain(A * 2, W, None, torch.bfloat16)
    assert _bits(y2, y1 * 2)
    X = torch.randn(M, C, device=DEV, generator=g).to(torch.bfloat16)
    w1, c1 = gemm_wgrad(y1, X)
    w2, c2 = gemm_wgrad(y1 * 2, X)
    assert _bits(w2, w1 * 2) and _bits(c2, c1 * 2)
    # fc2 + layer-scale + residual with gamma = 0 returns the shortcut untouched
    sc = torch.randn(M, C, device=DEV, generator=g)
    W2 = (torch.randn(C, 4 * C, device=DEV, generator=g) / (4 * C) ** 0.5).to(torch.bfloat16)
    out = gemm_scale_res(y1, W2, torch.zeros(C, device=DEV), torch.zeros(C, device=DEV), None, H * H, sc, torch.float32)
    assert _bits(out, sc)
    # GELU epilogue: g(h) - g(-h) = h for every element (GELU(x) - GELU(-x) = x), through two full-size launches
    _, ga = gemm_bias_gelu(A, W, torch.zeros(4 * C, device=DEV))
    _, gb = gemm_bias_gelu(-A, W, torch.zeros(4 * C, device=DEV))
    h = gemm_plain(A, W, None, torch.bfloat16).float()
    assert ((ga.float() - gb.float() - h).abs() <= 2e-2 * (1 + h.abs())).all()


@pytest.mark.parametrize("C,H", STAGES)
def test_dwconv_layernorm_full_size_identities(C, H):
    g = torch.Generator(device=DEV).manual_seed(H)
    x = torch.randn(N, H, H, C, device=DEV, generator=g)
    w = torch.randn(C, 1, 7, 7, device=DEV, generator=g) * 0.1
    b = torch.randn(C, device=DEV, generator=g) * 0.1
    ones, zeros = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
    M = N * H * H
    # fp32 path: linearity of the convolution in x, and LayerNorm rows with zero mean / unit variance
    y1, xn, mean, rstd = dwconv7_ln_fwd(x, w, b, ones, zeros, 1e-6, torch.float32)
    y2, _, _, _ = dwconv7_ln_fwd(2 * x, w, zeros, ones, zeros, 1e-6, torch.float32)
    assert max_rel(y2, 2 * (y1 - b)) <= 1e-6
    assert xn.mean(-1).abs().max().item() <= 1e-5 and (xn.var(-1, unbiased=False) - 1).abs().max().item() <= 1e-3
    assert max_rel(mean, y1.mean(-1)) <= 1e-5
    # adjoint identities: <conv(x) - b, dy> = <x, conv^T(dy)> = <w, dW>, and db = column sums of dy
    dy = torch.randn(M, C, device=DEV, generator=g)
    dx = dwconv7_dgrad(dy, w, None, (N, H, H, C), torch.float32)
    dw, db = dwconv7_wgrad(dy, x, P=max(1, 148 // (C // 32)))
    lhs = ((y1 - b).double() * dy.double()).sum().item()
    assert abs((x.double() * dx.double()).sum().item() - lhs) <= 1e-5 * abs(lhs) + 1e-3
    # (each dW entry is a sum over 10^5..10^6 pixels accumulated in fp32 and the 49*C terms cancel: tolerance on the terms' scale)
    wd = w.double() * dw.double()
    assert abs(wd.sum().item() - lhs) <= 2e-7 * wd.abs().sum().item() + 1e-3
    assert max_rel(db, dy.sum(0)) <= 1e-5
    # LayerNorm backward: dy_ln is orthogonal to the constant vector and to the normalised row
    dyl, dlw, dlb = ln_bwd(dy, y1, mean, rstd, ones, torch.float32, P=296)
    assert dyl.sum(-1).abs().max().item() <= 2e-3
    assert (dyl * (y1 - mean[:, None]) * rstd[:, None]).sum(-1).abs().max().item() <= 2e-2
    assert max_rel(dlb, dy.sum(0)) <= 1e-5
    # bf16 activations (the headline path): same launches, bf16 bar
    yb, xnb, _, _ = dwconv7_ln_fwd(x, w, b, ones, zeros, 1e-6, torch.bfloat16)
    assert max_rel(yb.float(), y1) <= 2e-2 and max_rel(xnb.float(), xn) <= 2e-2


def test_loss_and_label_mixing_on_the_simplex():
    B, K = 256, 1000
    g = torch.Generator(device=DEV).manual_seed(1)
    t = torch.randint(0, K, (B,), device=DEV, generator=g)
    for lam in (0.5, 0.3141592653589793):
        y = P.mixup_target(t, K, lam, 0.1)
        assert (y.sum(-1) - 1).abs().max().item() <= 2e-6 and (y >= 0).all()
        if lam == 0.5:
            assert _bits(y, y.flip(0))                       # symmetric mix: both members of a pair get the same target
    x = torch.randn(B, K, device=DEV, generator=g, requires_grad=True)
    loss = P.SoftTargetCrossEntropy()(x, y)
    loss.backward()
    assert loss.item() > 0 and x.grad.sum(-1).abs().max().item() <= 1e-6          # softmax - t sums to zero per row
    u = torch.zeros(B, K, device=DEV)
    assert abs(P.SoftTargetCrossEntropy()(u, y).item() - torch.log(torch.tensor(float(K))).item()) <= 1e-5


def test_image_mixing_symmetry_full_batch():
    g = torch.Generator(device=DEV).manual_seed(2)
    x = torch.randn(256, 3, 224, 224, device=DEV, generator=g)
    m = x.clone()
    ops.mixup_batch(m, 0.5)
    assert _bits(m, m.flip(0))                               # 0.5*a + 0.5*b in both orders
    keep = torch.empty_like(x)
    m2 = x.clone()
    ops.mixup_batch(m2, 1.0, original_out=keep)
    assert _bits(m2, x) and _bits(keep, x)                   # lam = 1: identity (x*1 + y*0), copy exact
    c = x.clone()
    ops.mixup_batch(c, 0.7, box=(10, 200, 30, 190))
    ops.mixup_batch(c, 0.7, box=(10, 200, 30, 190))
    assert _bits(c, x)                                       # the cutmix box swap is an involution


def test_ema_fixed_points_full_model():
    m = P.create_model("convnext_tiny", num_classes=1000).to(DEV)
    e = P.ModelEmaV3(m, decay=0.9995, device=torch.device(DEV))
    before = [p.clone() for p in e.module.parameters()]
    e.update(m)                                              # ema == model: lerp(e, p, w) with p == e is a fixed point
    assert all(_bits(a, b) for a, b in zip(before, e.module.parameters()))
    with torch.no_grad():
        for p in m.parameters():
            p.add_(1.0)
    e.update(m)
    w = 1.0 - 0.9995
    for a, b in zip(before, e.module.parameters()):          # e + w*((e+1) - e): the step is w up to one rounding of (p - e)
        assert ((b - a) - w).abs().max().item() <= 2e-7 + 1.2e-7 * a.abs().max().item()
    assert sum(p.numel() for p in m.parameters()) == 28_589_128
