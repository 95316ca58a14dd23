"""a1 dwconv7x7 + a2 channels-last LayerNorm, forward and backward kernels vs stock torch ops (fp32 reference).
fp32 bar <= 1e-4 relative; bf16 bar <= 2e-2 (BASELINE.json north_star)."""
import pytest
import torch
import torch.nn.functional as F

from cabi import (dwconv7_dgrad, dwconv7_dgrad_dz, dwconv7_ln_fwd, dwconv7_wgrad, grad_prep, ln_bwd, ln_fwd, max_rel)

pytestmark = pytest.mark.gpu
DEV = "cuda"

SHAPES = [(2, 56, 56, 96), (2, 28, 28, 192), (3, 14, 14, 384), (4, 7, 7, 768), (1, 9, 13, 32), (2, 1, 1, 64),
          (1, 12, 12, 1536), (1, 5, 40, 128)]


def _ref_fwd(x_nhwc, w, b, ln_w, ln_b, eps, act_dtype):
    x = x_nhwc.permute(0, 3, 1, 2).float()
    y = F.conv2d(x, w, b, padding=3, groups=x.shape[1]).permute(0, 2, 3, 1)
    y = y.to(act_dtype).float()                          # autocast: the conv output is rounded to the act dtype
    xn = F.layer_norm(y, (y.shape[-1],), ln_w, ln_b, eps)
    return y, xn


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("dtypes", [(torch.float32, torch.float32), (torch.float32, torch.bfloat16),
                                    (torch.bfloat16, torch.bfloat16)])
def test_dwconv_ln_fwd(shape, dtypes):
    sd, ad = dtypes
    N, H, W, C = shape
    g = torch.Generator().manual_seed(H * C)
    x = torch.randn(N, H, W, C, generator=g).to(sd).to(DEV)
    w = (torch.randn(C, 1, 7, 7, generator=g) * 0.2).to(DEV)
    b = torch.randn(C, generator=g).to(DEV) * 0.1
    lw = (1 + 0.1 * torch.randn(C, generator=g)).to(DEV)
    lb = (0.1 * torch.randn(C, generator=g)).to(DEV)
    y, xn, mean, rstd = dwconv7_ln_fwd(x, w, b, lw, lb, 1e-6, ad)
    yr, xnr = _ref_fwd(x, w, b, lw, lb, 1e-6, ad)
    tol = 1e-4 if ad == torch.float32 else 2e-2
    assert max_rel(y.float().view(N, H, W, C), yr) <= (2e-5 if ad == torch.float32 else 1e-2)
    assert max_rel(xn.float().view(N, H, W, C), xnr) <= tol
    yf = y.float()
    assert max_rel(mean, yf.mean(-1)) <= 1e-4 or (mean - yf.mean(-1)).abs().max() < 1e-5
    assert max_rel(rstd, (yf.var(-1, unbiased=False) + 1e-6).rsqrt()) <= 1e-4


@pytest.mark.parametrize("M,C", [(1000, 96), (37, 192), (513, 384), (64, 768), (5, 1024), (9, 1536), (3, 2048), (11, 40), (256, 3 * 4),
                                 (4099, 128), (777, 256), (300, 512), (20050, 96), (3001, 192), (1031, 384), (130, 56)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_ln_fwd_bwd(M, C, dt):
    g = torch.Generator().manual_seed(M + C)
    x = torch.randn(M, C, generator=g).to(dt).to(DEV)
    lw = (1 + 0.1 * torch.randn(C, generator=g)).to(DEV)
    lb = (0.1 * torch.randn(C, generator=g)).to(DEV)
    dout = torch.randn(M, C, generator=g).to(dt).to(DEV)
    out, mean, rstd = ln_fwd(x, lw, lb, 1e-6, torch.float32)
    xr = x.float().requires_grad_(True)
    lwr, lbr = lw.clone().requires_grad_(True), lb.clone().requires_grad_(True)
    ref = F.layer_norm(xr, (C,), lwr, lbr, 1e-6)
    assert max_rel(out, ref) <= 1e-5
    ref.backward(dout.float())
    dx, dw, db = ln_bwd(dout, x, mean, rstd, lw, dt, P=7)
    tol = 1e-4 if dt == torch.float32 else 1e-2
    assert max_rel(dx.float(), xr.grad) <= tol
    assert max_rel(dw, lwr.grad) <= 1e-4
    assert max_rel(db, lbr.grad) <= 1e-4


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("dtypes", [(torch.float32, torch.float32), (torch.float32, torch.bfloat16),
                                    (torch.bfloat16, torch.bfloat16)])
def test_dwconv_bwd(shape, dtypes):
    sd, ad = dtypes
    N, H, W, C = shape
    g = torch.Generator().manual_seed(H + C)
    x = torch.randn(N, H, W, C, generator=g).to(sd).to(DEV)
    w = (torch.randn(C, 1, 7, 7, generator=g) * 0.2).to(DEV)
    dy = torch.randn(N * H * W, C, generator=g).to(ad).to(DEV)
    dres = torch.randn(N, H, W, C, generator=g).to(sd).to(DEV)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    br = torch.zeros(C, device=DEV, requires_grad=True)
    F.conv2d(xr, wr, br, padding=3, groups=C).backward(dy.float().view(N, H, W, C).permute(0, 3, 1, 2))
    dx = dwconv7_dgrad(dy, w, dres, shape, sd)
    dx0 = dwconv7_dgrad(dy, w, None, shape, sd)
    tol = 1e-4 if sd == torch.float32 else 2e-2
    ref_dx = xr.grad.permute(0, 2, 3, 1)
    assert max_rel(dx0.float(), ref_dx) <= tol
    assert max_rel(dx.float(), ref_dx + dres.float()) <= tol
    dw, db = dwconv7_wgrad(dy, x, P=5)
    assert max_rel(dw, wr.grad) <= 1e-4
    assert max_rel(db, br.grad) <= 1e-4
    dw2, _ = dwconv7_wgrad(dy, x, P=64)                   # different CTA count -> same sums within fp32 reassociation
    assert max_rel(dw2, wr.grad) <= 1e-4


@pytest.mark.parametrize("shape", SHAPES + [(16, 56, 56, 96), (32, 14, 14, 384)])
@pytest.mark.parametrize("sd", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("with_dp", [True, False])
def test_dwconv_dgrad_with_upstream_operand_copy(shape, sd, with_dp):
    """cnx_dwconv7_dgrad_dz: dx is bit-identical to cnx_dwconv7_dgrad's and the folded copy dz_up to cnx_grad_prep(dx, dp_up)."""
    N, H, W, C = shape
    g = torch.Generator().manual_seed(H * 3 + C)
    w = (torch.randn(C, 1, 7, 7, generator=g) * 0.2).to(DEV)
    dy = torch.randn(N * H * W, C, generator=g).to(torch.bfloat16).to(DEV)
    dres = torch.randn(N, H, W, C, generator=g).to(sd).to(DEV)
    dp = ((torch.rand(N, generator=g) < 0.7).float() / 0.7).to(DEV) if with_dp else None
    dx_ref = dwconv7_dgrad(dy, w, dres, shape, sd)
    dx, dz = dwconv7_dgrad_dz(dy, w, dres, shape, sd, dp)
    assert torch.equal(dx, dx_ref)
    assert torch.equal(dz, grad_prep(dx_ref, dp, torch.bfloat16))


@pytest.mark.parametrize("shape", SHAPES)
def test_dwconv_ln_fwd_split_operand(shape):
    """cnx_dwconv7_ln_fwd_x3: the LayerNorm half writes [hi | mid | hi] of the fp32 xn that cnx_dwconv7_ln_fwd produces."""
    from cabi import dwconv7_ln_fwd_x3
    N, H, W, C = shape
    g = torch.Generator().manual_seed(N * H + C)
    x = torch.randn(N, H, W, C, generator=g).to(DEV)
    w = (0.1 * torch.randn(C, 1, 7, 7, generator=g)).to(DEV)
    b = (0.1 * torch.randn(C, generator=g)).to(DEV)
    lw = (1 + 0.1 * torch.randn(C, generator=g)).to(DEV)
    lb = (0.1 * torch.randn(C, generator=g)).to(DEV)
    y, xn, mean, rstd = dwconv7_ln_fwd(x, w, b, lw, lb, 1e-6, torch.float32)
    y3, a3, mean3, rstd3 = dwconv7_ln_fwd_x3(x, w, b, lw, lb, 1e-6)
    assert torch.equal(y, y3) and torch.equal(mean, mean3) and torch.equal(rstd, rstd3)
    hi, mid = a3[:, :C], a3[:, C:2 * C]
    assert torch.equal(hi, xn.to(torch.bfloat16)) and torch.equal(a3[:, 2 * C:], hi)
    assert torch.equal(mid, (xn - hi.float()).to(torch.bfloat16))
    # two-segment form [hi | mid] (row stride 2C) for a consumer whose K loop wraps
    y2, a2, mean2, rstd2 = dwconv7_ln_fwd_x3(x, w, b, lw, lb, 1e-6, 2)
    assert torch.equal(y2, y3) and torch.equal(a2, a3[:, :2 * C]) and torch.equal(mean2, mean3)
