"""stem (Conv2d 4x4 s4 + LayerNorm2d) and downsample (LayerNorm2d + Conv2d 2x2 s2) as patchify + GEMM +
LayerNorm kernels, against stock torch ops (fp32 reference; fp32 bar <= 1e-4 relative, bf16 bar <= 2e-2), plus the raw
re-ordering kernels (bit-exact: pure data movement)."""
import pytest
import torch
import torch.nn.functional as F

from cabi import max_rel
from imageclassification_b200 import _lib as L, ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _st():
    return L.stream()


@pytest.mark.parametrize("shape", [(2, 3, 32, 48), (1, 3, 224, 224), (3, 4, 8, 8)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_patchify4_is_unfold(shape, dt):
    N, C, H, W = shape
    x = torch.randn(shape, device=DEV)
    out = torch.empty((N * (H // 4) * (W // 4), C * 16), dtype=dt, device=DEV)
    L.check(L.load().cnx_patchify4_nchw(L.ptr(x), N, C, H, W, L.ptr(out), L.dt(dt), _st()), "patchify4")
    ref = F.unfold(x, kernel_size=4, stride=4).transpose(1, 2).reshape(-1, C * 16).to(dt)     # (ci, ky, kx) flattening
    assert torch.equal(out, ref)


@pytest.mark.parametrize("shape", [(2, 4, 6, 8), (1, 56, 56, 96), (3, 2, 2, 32)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_patch2_roundtrip(shape, dt):
    N, H, W, C = shape
    x = torch.randn(shape, device=DEV).to(dt)
    g = torch.empty((N, H // 2, W // 2, 4 * C), dtype=dt, device=DEV)
    L.check(L.load().cnx_patch2(L.ptr(x), L.dt(dt), N, H, W, C, L.ptr(g), 1, _st()), "patch2")
    ref = x.view(N, H // 2, 2, W // 2, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(N, H // 2, W // 2, 4 * C)
    assert torch.equal(g, ref)
    back = torch.empty_like(x)
    L.check(L.load().cnx_patch2(L.ptr(g), L.dt(dt), N, H, W, C, L.ptr(back), 0, _st()), "patch2")
    assert torch.equal(back, x)


def _run(fn, params, x, dout, autocast):
    for p in params:
        p.grad = None
    xr = x.clone().requires_grad_(x.is_floating_point() and fn.__name__ != "stem")
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        out = fn(xr)
    out.float().backward(dout)
    return out.float(), [p.grad.clone() for p in params], (xr.grad.clone() if xr.grad is not None else None)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("N,H,C", [(2, 32, 96), (1, 64, 64), (3, 16, 32)])
def test_stem_fwd_bwd(mode, N, H, C):
    torch.manual_seed(N * H + C)
    conv = torch.nn.Conv2d(3, C, 4, 4).to(DEV)
    lw = (1 + 0.1 * torch.randn(C)).to(DEV).requires_grad_(True)
    lb = (0.1 * torch.randn(C)).to(DEV).requires_grad_(True)
    x = torch.randn(N, 3, H, H, device=DEV)
    dout = torch.randn(N, C, H // 4, H // 4, device=DEV)
    params = [conv.weight, conv.bias, lw, lb]

    def ref(xx):
        y = conv(xx)
        return F.layer_norm(y.permute(0, 2, 3, 1), (C,), lw, lb, 1e-6).permute(0, 3, 1, 2)

    def stem(xx):
        return ops.stem_forward(xx, conv.weight, conv.bias, lw, lb, 1e-6)

    ro, rg, _ = _run(ref, params, x, dout, mode == "bf16")
    oo, og, _ = _run(stem, params, x, dout, mode == "bf16")
    tol = 1e-4 if mode == "fp32" else 2e-2
    assert max_rel(oo, ro) <= tol
    for a, b in zip(og, rg):
        assert max_rel(a, b) <= (2e-4 if mode == "fp32" else 2e-2)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("N,H,C", [(2, 28, 96), (1, 14, 192), (3, 8, 32), (2, 14, 384)])
def test_downsample_fwd_bwd(mode, N, H, C):
    torch.manual_seed(N * H + C + 1)
    conv = torch.nn.Conv2d(C, 2 * C, 2, 2).to(DEV)
    lw = (1 + 0.1 * torch.randn(C)).to(DEV).requires_grad_(True)
    lb = (0.1 * torch.randn(C)).to(DEV).requires_grad_(True)
    x = torch.randn(N, C, H, H, device=DEV).contiguous(memory_format=torch.channels_last)
    dout = torch.randn(N, 2 * C, H // 2, H // 2, device=DEV)
    params = [conv.weight, conv.bias, lw, lb]

    def ref(xx):
        y = F.layer_norm(xx.permute(0, 2, 3, 1), (C,), lw, lb, 1e-6).permute(0, 3, 1, 2)
        return conv(y)

    def down(xx):
        return ops.downsample_forward(xx, lw, lb, conv.weight, conv.bias, 1e-6)

    ro, rg, rx = _run(ref, params, x, dout, mode == "bf16")
    oo, og, ox = _run(down, params, x, dout, mode == "bf16")
    tol = 1e-4 if mode == "fp32" else 2e-2
    assert max_rel(oo, ro) <= tol
    assert max_rel(ox.float(), rx.float()) <= (2e-4 if mode == "fp32" else 3e-2)
    for a, b in zip(og, rg):
        assert max_rel(a, b) <= (2e-4 if mode == "fp32" else 3e-2)


def test_weight_prep_registry_refreshes_all_in_one_launch():
    """The derived-weight registry rebuilds every stale layout with ONE cnx_weight_prep_multi call."""
    ws = [torch.randn(64 + 32 * i, 96, device=DEV) for i in range(3)]
    sc = torch.rand(64, device=DEV) + 0.5
    outs = [ops._weight_prep(w, 1, None, torch.bfloat16) for w in ws] + [ops._weight_prep(ws[0], 2, sc, torch.bfloat16)]
    for w, o in zip(ws, outs):
        assert torch.equal(o, w.t().to(torch.bfloat16))
    assert torch.equal(outs[3], (ws[0] * sc[:, None]).t().to(torch.bfloat16))
    with torch.no_grad():
        for w in ws:
            w.mul_(2.0)                        # bumps the versions: everything is stale
    c0 = L.CALL_COUNTS["cnx_weight_prep_multi"]
    again = [ops._weight_prep(w, 1, None, torch.bfloat16) for w in ws] + [ops._weight_prep(ws[0], 2, sc, torch.bfloat16)]
    assert L.CALL_COUNTS["cnx_weight_prep_multi"] - c0 == 1
    for w, o in zip(ws, again):
        assert torch.equal(o, w.t().to(torch.bfloat16))
    assert torch.equal(again[3], (ws[0] * sc[:, None]).t().to(torch.bfloat16))
    # five layouts of one source (what a Block's fc weight has) share its tiles: still one launch, every layout right
    w = ws[1]
    lay = lambda: [ops._weight_prep(w, 0, None, torch.bfloat16), ops._weight_prep(w, 3, None, torch.bfloat16),
                   ops._weight_prep(w, 1, None, torch.float32), ops._weight_prep(w, 1, None, torch.bfloat16)]
    lay()
    with torch.no_grad():
        w.add_(0.125)
    c0 = L.CALL_COUNTS["cnx_weight_prep_multi"]
    o0, o3, o1f, o1 = lay()
    assert L.CALL_COUNTS["cnx_weight_prep_multi"] - c0 == 1
    hi = w.to(torch.bfloat16)
    assert torch.equal(o0, hi) and torch.equal(o1f, w.t()) and torch.equal(o1, w.t().to(torch.bfloat16))
    assert torch.equal(o3, torch.cat([hi, hi, (w - hi.float()).to(torch.bfloat16)], dim=1))
    assert torch.equal(ops._weight_prep(ws[0], 2, sc, torch.bfloat16), (ws[0] * sc[:, None]).t().to(torch.bfloat16))


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("N,C,H,K", [(8, 768, 7, 1000), (256, 768, 7, 1000), (5, 96, 3, 16), (4, 1536, 12, 1000), (3, 128, 1, 8)])
def test_head_fwd_bwd(mode, N, C, H, K):
    """NormMlpClassifierHead (pool -> LayerNorm2d -> fc) on libcnx kernels vs the oracle head: logits, input and all parameter
    gradients.  K not a multiple of 8 (e.g. the 2-class config) keeps the ATen modules and is covered by the model tests."""
    from imageclassification_b200 import modules as PM
    from oracle import convnext as OC
    torch.manual_seed(N + C + K)
    o = OC.Head(C, K).to(DEV)
    p = PM.NormMlpClassifierHead(C, K).to(DEV)
    with torch.no_grad():
        o.norm.weight.add_(0.1 * torch.randn_like(o.norm.weight))
        o.norm.bias.normal_(0, 0.1)
        o.fc.weight.normal_(0, 0.05)
        o.fc.bias.normal_(0, 0.1)
    p.load_state_dict(o.state_dict())
    x = torch.randn(N, C, H, H, device=DEV).contiguous(memory_format=torch.channels_last)
    dout = torch.randn(N, K, device=DEV)
    res = []
    for m in (o, p):
        xi = x.clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
            y = m(xi)
        y.backward(dout.to(y.dtype))
        res.append((y, xi.grad, {n: q.grad for n, q in m.named_parameters()}))
    (yo, dxo, go), (yp, dxp, gp) = res
    tol = 1e-4 if mode == "fp32" else 2e-2
    assert yp.dtype == yo.dtype and yp.shape == yo.shape
    assert max_rel(yp, yo) <= tol
    assert max_rel(dxp, dxo) <= tol
    for n in go:
        assert max_rel(gp[n], go[n]) <= tol, n
    with torch.no_grad():                                    # fp32 no-grad: split-operand tensor-core GEMM
        assert max_rel(p(x), o(x)) <= 1e-4


def test_downsample_widened_output_and_gradient_handoff(monkeypatch):
    """Under bf16 autocast the downsample conv of a stage writes the fp32 tensor its first Block would widen the bf16 output to
    (CNX_GEMM_OUT_ROUND_BF16: same values, no cast pass), and that Block's backward hands back the bf16 copy of its dx (no cast
    pass either).  Outputs and every gradient are bit-identical to the path with the bf16 output and the two ATen casts."""
    from imageclassification_b200 import modules as PM
    torch.manual_seed(4)
    stage = PM.ConvNeXtStage(96, 192, stride=2, depth=2, drop_path_rates=[0.0, 0.25], ls_init_value=1.0).to(DEV)
    with torch.no_grad():
        for p in stage.parameters():
            if p.ndim == 1:
                p.add_(0.05 * torch.randn_like(p))
    g = torch.Generator().manual_seed(2)
    x = torch.randn(6, 96, 28, 28, generator=g).to(DEV)
    dout = torch.randn(6, 192, 14, 14, generator=g).to(DEV)

    def run(widen):
        stage.zero_grad()
        xi = x.clone().requires_grad_(True)
        torch.manual_seed(21)
        c0 = dict(L.CALL_COUNTS)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            if widen:
                y = stage(xi)
            else:
                ds = stage.downsample
                t = torch.ops.cnx.downsample_forward(xi, ds[0].weight, ds[0].bias, ds[1].weight, ds[1].bias, ds[0].eps, False)
                assert t.dtype == torch.bfloat16
                y = stage.blocks(t)
        y.backward(dout)
        n = {k: L.CALL_COUNTS[k] - c0.get(k, 0) for k in ("cnx_grad_prep", "cnx_dwconv7_dgrad_dz")}
        return y, xi.grad, [p.grad.clone() for p in stage.parameters()], n

    y1, dx1, g1, n1 = run(True)
    assert n1 == {"cnx_grad_prep": 1, "cnx_dwconv7_dgrad_dz": 2}, n1          # both Blocks hand their dx copy upstream
    y0, dx0, g0, n0 = run(False)
    assert n0 == {"cnx_grad_prep": 1, "cnx_dwconv7_dgrad_dz": 1}, n0
    assert y1.dtype == torch.float32 and torch.equal(y1, y0) and torch.equal(dx1, dx0)
    for a, b in zip(g1, g0):
        assert torch.equal(a, b)
    monkeypatch.setattr(ops, "DZ_HANDOFF", False)
    y2, dx2, g2, n2 = run(True)
    assert n2 == {"cnx_grad_prep": 2, "cnx_dwconv7_dgrad_dz": 0}, n2
    assert torch.equal(y2, y0) and torch.equal(dx2, dx0)
    for a, b in zip(g2, g0):
        assert torch.equal(a, b)
    # the widened tensor itself: fp32 holding bf16-representable values, equal to the bf16 output
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        ds = stage.downsample
        a = torch.ops.cnx.downsample_forward(x, ds[0].weight, ds[0].bias, ds[1].weight, ds[1].bias, ds[0].eps, True)
        b = torch.ops.cnx.downsample_forward(x, ds[0].weight, ds[0].bias, ds[1].weight, ds[1].bias, ds[0].eps, False)
    assert a.dtype == torch.float32 and b.dtype == torch.bfloat16 and torch.equal(a, b.float())
