"""ConvNeXt Block fwd+bwd (BASELINE config 5 sweep C x H) and the full ConvNeXt-T:
libcnx modules vs the oracle modules with identical weights and inputs.
Bars: fp32 <= 1e-4 relative on outputs and gradients; bf16 autocast <= 2e-2 (BASELINE.json north_star)."""
import pytest
import torch

import imageclassification_b200 as P
from cabi import max_rel
from oracle import convnext as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _pair_block(C, drop_path, gamma_init, seed):
    torch.manual_seed(seed)
    o = O.ConvNeXtBlock(C, drop_path=drop_path, ls_init_value=gamma_init).to(DEV)
    with torch.no_grad():                                   # non-trivial biases / affine params / gamma
        for n, p in o.named_parameters():
            if n.endswith("bias"):
                p.normal_(0, 0.05)
            elif n == "norm.weight":
                p.add_(0.1 * torch.randn_like(p))
            elif n == "gamma":
                p.mul_(1 + 0.2 * torch.randn_like(p))
    p_ = P.ConvNeXtBlock(C, drop_path=drop_path, ls_init_value=gamma_init).to(DEV)
    p_.load_state_dict(o.state_dict())
    return o, p_


def _run(mod, x, dout, autocast, seed):
    x = x.clone().requires_grad_(True)
    torch.manual_seed(seed)                                 # same drop-path mask draw in both
    if autocast:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = mod(x)
    else:
        y = mod(x)
    y.backward(dout.to(y.dtype))
    return y, x.grad, {n: p.grad for n, p in mod.named_parameters()}


SWEEP = [(96, 56), (192, 28), (384, 14), (768, 7), (96, 7), (768, 14), (128, 9), (1536, 12)]


@pytest.mark.parametrize("C,H", SWEEP)
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("drop_path", [0.0, 0.4])
def test_block_fwd_bwd(C, H, mode, drop_path):
    N = 4 if C * H * H > 200000 else 6
    o, p = _pair_block(C, drop_path, 1.0, C + H)
    g = torch.Generator().manual_seed(H)
    x = torch.randn(N, C, H, H, generator=g).to(DEV)
    dout = torch.randn(N, C, H, H, generator=g).to(DEV)
    ac = mode == "bf16"
    yo, dxo, go = _run(o, x, dout, ac, 123)
    yp, dxp, gp = _run(p, x, dout, ac, 123)
    assert yp.shape == yo.shape and yp.dtype == yo.dtype == torch.float32    # residual stream stays fp32 under autocast
    tol = 1e-4 if mode == "fp32" else 2e-2
    assert max_rel(yp, yo) <= tol
    assert max_rel(dxp, dxo) <= tol
    for n in go:
        assert gp[n] is not None, n
        assert gp[n].shape == go[n].shape
        assert max_rel(gp[n], go[n]) <= tol, n


def test_block_tiny_gamma_and_eval():
    """gamma = 1e-6 (the timm default): the branch nearly vanishes from y, but dgamma must still be right."""
    o, p = _pair_block(96, 0.0, 1e-6, 9)
    x = torch.randn(2, 96, 14, 14, device=DEV)
    dout = torch.randn_like(x)
    yo, dxo, go = _run(o, x, dout, False, 1)
    yp, dxp, gp = _run(p, x, dout, False, 1)
    assert max_rel(yp, yo) <= 1e-6
    assert max_rel(gp["gamma"], go["gamma"]) <= 1e-4
    assert max_rel(dxp, dxo) <= 1e-6
    o.eval(); p.eval()
    with torch.no_grad():
        assert max_rel(p(x), o(x)) <= 1e-6
    o2, p2 = _pair_block(96, 0.5, 1.0, 9)
    o2.eval(); p2.eval()                                     # drop-path is the identity in eval mode
    with torch.no_grad():
        assert max_rel(p2(x), o2(x)) <= 1e-4


def test_block_channels_last_input_and_errors():
    o, p = _pair_block(64, 0.0, 1.0, 4)
    x = torch.randn(2, 64, 9, 11, device=DEV).contiguous(memory_format=torch.channels_last)
    assert max_rel(p(x), o(x)) <= 1e-4
    with pytest.raises(RuntimeError):
        P.ConvNeXtBlock(64)(torch.randn(1, 64, 8, 8))        # CPU tensors: no fallback, loud failure
    with pytest.raises(NotImplementedError):
        P.ConvNeXtBlock(64, kernel_size=3)


def _pair_model(name, num_classes, dpr, gamma_init, seed):
    torch.manual_seed(seed)
    o = O.create_model(name, num_classes=num_classes, drop_path_rate=dpr, ls_init_value=gamma_init).to(DEV)
    p = P.create_model(name, num_classes=num_classes, drop_path_rate=dpr, ls_init_value=gamma_init).to(DEV)
    p.load_state_dict(o.state_dict())
    return o, p


@pytest.mark.parametrize("mode,gamma_init,dpr", [("fp32", 1.0, 0.0), ("fp32", 1e-6, 0.05), ("bf16", 1.0, 0.0), ("bf16", 1e-6, 0.05)])
def test_convnext_tiny_fwd_bwd(mode, gamma_init, dpr):
    """BASELINE config 1 shape (batch 8, 2 classes, 224^2): logits and ALL 182 gradients vs the oracle."""
    o, p = _pair_model("convnext_tiny", 2, dpr, gamma_init, 88)
    g = torch.Generator().manual_seed(88)
    x = torch.randn(8, 3, 224, 224, generator=g).to(DEV)
    t = torch.softmax(torch.randn(8, 2, generator=g), -1).to(DEV)
    outs = []
    for m in (o, p):
        torch.manual_seed(7)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
            logits = m(x)
            loss = torch.sum(-t * torch.log_softmax(logits.float(), -1), -1).mean()
        loss.backward()
        outs.append((logits.float(), {n: q.grad for n, q in m.named_parameters()}))
    (lo, go), (lp, gp) = outs
    tol = 1e-4 if mode == "fp32" else 2e-2
    assert max_rel(lp, lo) <= tol
    worst = max((max_rel(gp[n], go[n]), n) for n in go)
    # gradients: per-tensor max-relative error; bf16 errors accumulate over 18 blocks, bar is per north_star 2e-2
    assert worst[0] <= (1e-3 if mode == "fp32" else 5e-2), worst
    import math
    tot_p = math.sqrt(sum((gp[n].double() ** 2).sum().item() for n in go))
    tot_d = math.sqrt(sum(((gp[n].double() - go[n].double()) ** 2).sum().item() for n in go))
    assert tot_d / tot_p <= tol, (tot_d / tot_p)


@pytest.mark.parametrize("name,img,batch", [("convnext_base", 64, 4), ("convnext_large", 96, 2)])
def test_convnext_base_large_fwd_bwd_bf16(name, img, batch):
    """BASELINE configs 3 and 4 model families (widths 128..1024 and 192..1536, depths 3/3/27/3) at a reduced resolution:
    logits and the global gradient vector vs the oracle under bf16 autocast, mixup-style soft targets."""
    import math
    o, p = _pair_model(name, 1000, 0.1, 1.0, 5)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(batch, 3, img, img, generator=g).to(DEV)
    t = torch.softmax(torch.randn(batch, 1000, generator=g), -1).to(DEV)
    outs = []
    for m in (o, p):
        torch.manual_seed(7)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = m(x)
            loss = torch.sum(-t * torch.log_softmax(logits.float(), -1), -1).mean()
        loss.backward()
        outs.append((logits.float(), {n: q.grad for n, q in m.named_parameters()}))
    (lo, go), (lp, gp) = outs
    assert max_rel(lp, lo) <= 2e-2
    tot_p = math.sqrt(sum((gp[n].double() ** 2).sum().item() for n in go))
    tot_d = math.sqrt(sum(((gp[n].double() - go[n].double()) ** 2).sum().item() for n in go))
    assert tot_d / tot_p <= 2e-2, (tot_d / tot_p)


@pytest.mark.parametrize("C,H", SWEEP + [(64, 5)])
def test_block_fp32_nograd_forward_on_tensor_cores(C, H):
    """fp32 no-grad forward (the reference's accuracy forward / evaluate run outside autocast): split-operand tcgen05 GEMMs
    (include/cnx.h "x3") vs the oracle in fp32 — the fp32 bar (1e-4) — and vs this package's CUDA-core fp32 path."""
    from imageclassification_b200 import ops
    o, p = _pair_block(C, 0.0, 1.0, 3 * C + H)
    N = 3 if C * H * H < 200000 else 2
    x = torch.randn(N, C, H, H, device=DEV)
    with torch.no_grad():
        yo = o(x)
        assert ops.X3_FWD
        yp = p(x)
        ops.X3_FWD = False
        try:
            ys = p(x)
        finally:
            ops.X3_FWD = True
    assert yp.dtype == torch.float32
    assert max_rel(yp, yo) <= 1e-4
    assert max_rel(yp, ys) <= 1e-4
    # the branch itself (what the GEMMs produce), not hidden behind the shortcut
    assert max_rel(yp - x, yo - x) <= 1e-4


def test_convnext_tiny_fp32_nograd_forward_on_tensor_cores():
    o, p = _pair_model("convnext_tiny", 10, 0.0, 1.0, 17)
    x = torch.randn(4, 3, 224, 224, device=DEV)
    o.eval()
    p.eval()
    with torch.no_grad():
        lo, lp = o(x), p(x)
    assert max_rel(lp, lo) <= 1e-4


@pytest.mark.parametrize("C,H", [(96, 28), (192, 14)])
def test_block_backward_with_recomputed_gelu_prime(C, H, monkeypatch):
    """ops.RECOMPUTE_MAX_C (CNX_RECOMPUTE_C): the forward stores g only and the backward kernel recomputes GELU'(h) — the same
    gradients, bit for bit, as the default path that saves GELU'(h)."""
    from imageclassification_b200 import ops
    o, p = _pair_block(C, 0.0, 1.0, 7)
    g = torch.Generator().manual_seed(C)
    x = torch.randn(8, C, H, H, generator=g).to(DEV)
    dout = torch.randn(8, C, H, H, generator=g).to(DEV)
    y0, dx0, g0 = _run(p, x, dout, True, 1)
    p.zero_grad()
    monkeypatch.setattr(ops, "RECOMPUTE_MAX_C", 192)
    y1, dx1, g1 = _run(p, x, dout, True, 1)
    assert torch.equal(y0, y1) and torch.equal(dx0, dx1)
    for n in g0:
        assert torch.equal(g0[n], g1[n]), n
    yo, dxo, go = _run(o, x, dout, True, 1)
    assert max_rel(dx1, dxo) <= 2e-2


@pytest.mark.parametrize("drop_path", [0.0, 0.3])
def test_block_chain_backward_operand_handoff(drop_path, monkeypatch):
    """ops.DZ_HANDOFF: between consecutive Blocks under bf16 autocast the dwconv backward-data kernel of Block i also writes the
    bf16 operand copy of its dx that Block i-1's backward needs (instead of a cnx_grad_prep pass there).  Same gradients, bit for
    bit, as without the hand-off; the copy is only taken when the incoming gradient IS that dx (a tensor hook that replaces it
    falls back to cnx_grad_prep)."""
    from imageclassification_b200 import _lib as L, ops
    torch.manual_seed(3)
    C, H, N = 96, 28, 6
    blocks = torch.nn.Sequential(*[P.ConvNeXtBlock(C, drop_path=drop_path, ls_init_value=1.0) for _ in range(3)]).to(DEV)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, C, H, H, generator=g).to(DEV)
    dout = torch.randn(N, C, H, H, generator=g).to(DEV)

    def run(hook=False):
        blocks.zero_grad()
        xi = x.clone().requires_grad_(True)
        torch.manual_seed(11)
        c0 = dict(L.CALL_COUNTS)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            t = blocks[0](xi)
            t = blocks[1](t)
            if hook:
                t.register_hook(lambda gr: gr * 1.0)            # a NEW tensor reaches Block 1's backward
            y = blocks[2](t)
        y.backward(dout)
        n = {k: L.CALL_COUNTS[k] - c0.get(k, 0) for k in ("cnx_grad_prep", "cnx_dwconv7_dgrad_dz", "cnx_dwconv7_dgrad")}
        return y, xi.grad, [p.grad.clone() for p in blocks.parameters()], n

    y1, dx1, g1, n1 = run()
    assert n1 == {"cnx_grad_prep": 1, "cnx_dwconv7_dgrad_dz": 2, "cnx_dwconv7_dgrad": 1}, n1
    y2, dx2, g2, n2 = run(hook=True)
    assert n2["cnx_grad_prep"] == 2 and n2["cnx_dwconv7_dgrad_dz"] == 2, n2
    monkeypatch.setattr(ops, "DZ_HANDOFF", False)
    y0, dx0, g0, n0 = run()
    assert n0 == {"cnx_grad_prep": 3, "cnx_dwconv7_dgrad_dz": 0, "cnx_dwconv7_dgrad": 3}, n0
    for a, b in ((y1, y0), (dx1, dx0), (y2, y0), (dx2, dx0)):
        assert torch.equal(a, b)
    for a, b, c in zip(g1, g0, g2):
        assert torch.equal(a, b) and torch.equal(c, b)


@pytest.mark.parametrize("N,H,W", [(8, 56, 56), (3, 28, 28), (1, 9, 13), (5, 7, 7), (2, 16, 8)])
@pytest.mark.parametrize("drop_path,gamma_init", [(0.0, 1.0), (0.35, 1e-6), (0.0, None)])
def test_block_fused_x3_mlp_matches_unfused(N, H, W, drop_path, gamma_init, monkeypatch):
    """cnx_mlp_fused_fwd_x3 (C = 96: fc1 -> GELU -> split -> fc2 in one kernel, hidden activation on chip) against the unfused
    split-operand pair and against the fp32 oracle: whole and ragged 128-row tiles (M = 117 ... 25088), drop-path in train mode,
    tiny and absent layer scale.  Same products, another summation order: 2e-5 between the two; the fp32 bar (1e-4) vs the oracle."""
    from imageclassification_b200 import _lib as L, ops
    C = 96
    o, p = _pair_block(C, drop_path, gamma_init, N * H + W)
    x = torch.randn(N, C, H, W, device=DEV)
    if N * H * W < 128:
        pytest.skip("below one tile the Block takes the unfused pair")

    def run(mod):
        torch.manual_seed(5)                                    # the same drop-path draw
        with torch.no_grad():
            return mod(x)

    o.train()
    p.train()
    yo = run(o)
    c0 = L.CALL_COUNTS["cnx_mlp_fused_fwd_x3"]
    yf = run(p)
    assert L.CALL_COUNTS["cnx_mlp_fused_fwd_x3"] == c0 + 1
    monkeypatch.setattr(ops, "FUSED_MLP_X3", False)
    yu = run(p)
    assert L.CALL_COUNTS["cnx_mlp_fused_fwd_x3"] == c0 + 1
    assert yf.dtype == torch.float32
    assert max_rel(yf, yu) <= 2e-6 and max_rel(yf, yo) <= 1e-4
    if gamma_init is None or gamma_init >= 1e-2:          # with a 1e-6 layer scale the branch is below the ulp of x + branch
        assert max_rel(yf - x, yu - x) <= 2e-5, max_rel(yf - x, yu - x)
        assert max_rel(yf - x, yo - x) <= 1e-4, max_rel(yf - x, yo - x)
    else:                                                 # ... so look at the branch through a unit layer scale instead
        g0 = p.gamma.detach().clone()
        with torch.no_grad():
            p.gamma.fill_(1.0)
        monkeypatch.setattr(ops, "FUSED_MLP_X3", True)
        b1_ = run(p) - x
        monkeypatch.setattr(ops, "FUSED_MLP_X3", False)
        b0_ = run(p) - x
        with torch.no_grad():
            p.gamma.copy_(g0)
        assert max_rel(b1_, b0_) <= 2e-5, max_rel(b1_, b0_)
