"""a3..a8: the MLP GEMMs with fused epilogues.  fp32 CUDA-core path vs torch fp32 (<= 1e-4 relative);
bf16 tcgen05 path vs (i) the CUDA-core kernel on the same bf16 operands (tight: same operands, fp32 accumulate)
and (ii) torch fp32 on the bf16-rounded operands (<= 2e-2, the bf16 bar of BASELINE.json north_star)."""
import math

import pytest
import torch
import torch.nn.functional as F

from cabi import gemm_bias_gelu, gemm_dgelu, gemm_plain, gemm_scale_res, gemm_wgrad, max_rel
from imageclassification_b200 import _lib as L

pytestmark = pytest.mark.gpu
DEV = "cuda"
SIMT = L.CNX_GEMM_FORCE_SIMT

# (M, C): fc1 is [M,C]x[4C,C]^T, fc2 is [M,4C]x[C,4C]^T
MC = [(3136, 96), (784 * 2, 192), (196 * 3, 384), (49 * 4, 768), (100, 96), (1, 96), (129, 128), (300, 1536), (257, 32)]


def _mk(M, C, dt, seed):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(M, C, generator=g).to(dt).to(DEV)
    W1 = (torch.randn(4 * C, C, generator=g) / math.sqrt(C)).to(dt).to(DEV)
    b1 = (0.1 * torch.randn(4 * C, generator=g)).to(DEV)
    W2 = (torch.randn(C, 4 * C, generator=g) / math.sqrt(4 * C)).to(dt).to(DEV)
    b2 = (0.1 * torch.randn(C, generator=g)).to(DEV)
    gamma = torch.rand(C, generator=g).to(DEV) + 0.5
    return A, W1, b1, W2, b2, gamma


@pytest.mark.parametrize("M,C", MC)
@pytest.mark.parametrize("mode", ["f32", "bf16_simt", "bf16_tc"])
def test_mlp_forward_gemms(M, C, mode):
    dt = torch.float32 if mode == "f32" else torch.bfloat16
    flags = SIMT if mode == "bf16_simt" else 0
    A, W1, b1, W2, b2, gamma = _mk(M, C, dt, M + C)
    gp, gl = gemm_bias_gelu(A, W1, b1, flags)
    href = (A.float() @ W1.float().t() + b1).to(dt).float()   # h rounded to the activation dtype, as autocast's Linear output
    tol = 1e-4 if dt == torch.float32 else 1e-2
    assert max_rel(gl.float(), F.gelu(href)) <= tol
    hr = href.clone().requires_grad_(True)
    F.gelu(hr).sum().backward()
    assert max_rel(gp.float(), hr.grad) <= tol                # saved for backward: GELU'(h), not h
    rps = 7
    n_s = (M + rps - 1) // rps
    dp = (torch.rand(n_s, device=DEV) > 0.3).float() / 0.7
    for sdt in ([torch.float32] if dt == torch.float32 else [torch.float32, torch.bfloat16]):
        sc = torch.randn(M, C, device=DEV).to(sdt)
        out = gemm_scale_res(gl, W2, b2, gamma, dp, rps, sc, sdt, flags)
        z = gl.float() @ W2.float().t() + b2
        ref = sc.float() + dp.repeat_interleave(rps)[:M, None] * (gamma * z)
        assert max_rel(out.float(), ref) <= (1e-4 if dt == torch.float32 else 2e-2)
        out2 = gemm_scale_res(gl, W2, None, None, None, 1, None, sdt, flags)
        assert max_rel(out2.float(), gl.float() @ W2.float().t()) <= (1e-4 if dt == torch.float32 else 2e-2)


@pytest.mark.parametrize("M,C", MC)
@pytest.mark.parametrize("mode", ["f32", "bf16_simt", "bf16_tc"])
def test_mlp_backward_gemms(M, C, mode):
    dt = torch.float32 if mode == "f32" else torch.bfloat16
    flags = SIMT if mode == "bf16_simt" else 0
    A, W1, b1, W2, b2, gamma = _mk(M, C, dt, 2 * M + C)
    g = torch.Generator().manual_seed(5)
    dz = torch.randn(M, C, generator=g).to(dt).to(DEV)
    h = torch.randn(M, 4 * C, generator=g).to(dt).to(DEV)
    tol = 1e-4 if dt == torch.float32 else 2e-2
    # dgrad fc2 with GELU':  Bt = W2^T [4C, C]; the kernel takes gp = GELU'(h) as saved by the forward
    Bt = W2.t().contiguous()
    hf = h.float().requires_grad_(True)
    F.gelu(hf).sum().backward()
    gp = hf.grad.to(dt)
    dh = gemm_dgelu(dz, Bt, gp, flags)
    assert max_rel(dh.float(), (dz.float() @ W2.float()) * gp.float()) <= tol
    # dgrad fc1: dxn = dh . W1  (B = W1^T [C, 4C])
    dxn = gemm_plain(dh, W1.t().contiguous(), None, dt, flags)
    assert max_rel(dxn.float(), dh.float() @ W1.float()) <= tol
    if dt == torch.bfloat16:
        out32 = gemm_plain(dh, W1.t().contiguous(), b2, torch.float32, flags)
        assert max_rel(out32, dh.float() @ W1.float() + b2) <= 1e-3
    # wgrads + bias grads
    G, s = gemm_wgrad(dz, h, flags)                       # [C, 4C]
    assert max_rel(G, dz.float().t() @ h.float()) <= (1e-4 if dt == torch.float32 else 1e-3)
    assert max_rel(s, dz.float().sum(0)) <= 1e-4
    G1, s1 = gemm_wgrad(h, A, flags)                      # [4C, C]
    assert max_rel(G1, h.float().t() @ A.float()) <= (1e-4 if dt == torch.float32 else 1e-3)
    assert max_rel(s1, h.float().sum(0)) <= 1e-4


@pytest.mark.parametrize("M,C", [(3136, 96), (1000, 192), (260, 768)])
def test_tc_matches_simt_on_same_operands(M, C):
    """Same bf16 operands, both accumulate in fp32: differences are summation-order only."""
    A, W1, b1, W2, b2, gamma = _mk(M, C, torch.bfloat16, 3 * M + C)
    h0, g0 = gemm_bias_gelu(A, W1, b1, SIMT)
    h1, g1 = gemm_bias_gelu(A, W1, b1, 0)
    assert max_rel(h1.float(), h0.float()) <= 8e-3        # GELU'(h): <= 1 bf16 ulp on rounding flips (polynomial vs erff)
    assert max_rel(g1.float(), g0.float()) <= 8e-3
    o0 = gemm_scale_res(g0, W2, b2, gamma, None, 1, None, torch.float32, SIMT)
    o1 = gemm_scale_res(g0, W2, b2, gamma, None, 1, None, torch.float32, 0)
    assert max_rel(o1, o0) <= 1e-5
    G0, s0 = gemm_wgrad(g0, A, SIMT)
    G1, s1 = gemm_wgrad(g0, A, 0)
    assert max_rel(G1, G0) <= 1e-4 and max_rel(s1, s0) <= 1e-4


def test_gemm_argument_errors():
    lib = L.load()
    a = torch.zeros(8, 12, device=DEV)
    rc = lib.cnx_gemm_plain(L.ptr(a), L.ptr(a), None, L.ptr(a), 0, 8, 12, 12, 0, 0, L.stream())   # N % 8 != 0
    assert rc == -2 and b"multiple of 8" in lib.cnx_last_error_string()
    rc = lib.cnx_gemm_plain(None, L.ptr(a), None, L.ptr(a), 0, 8, 8, 12, 0, 0, L.stream())
    assert rc == -1


@pytest.mark.parametrize("M,C", [(3136, 96), (128, 96), (1000, 192), (777, 128), (50176, 96), (4 * 784 + 5, 192)])
@pytest.mark.parametrize("with_dp", [False, True])
def test_fused_mlp_forward_matches_two_gemm_path(M, C, with_dp):
    """The fused no-grad MLP kernel (hidden activation kept on chip) vs the fc1+GELU and fc2+residual GEMM kernels on the same
    bf16 operands: same roundings (h and g to bf16), fp32 accumulation — differences are summation order only."""
    from cabi import mlp_fused_fwd
    A, W1, b1, W2, b2, gamma = _mk(M, C, torch.bfloat16, 5 * M + C)
    rps = 49
    n_s = (M + rps - 1) // rps
    dp = ((torch.rand(n_s, device=DEV) > 0.3).float() / 0.7) if with_dp else None
    sc = torch.randn(M, C, device=DEV)
    _, g = gemm_bias_gelu(A, W1, b1, 0)
    ref = gemm_scale_res(g, W2, b2, gamma, dp, rps, sc, torch.float32, 0)
    out = mlp_fused_fwd(A, W1, b1, W2, b2, gamma, dp, rps, sc)
    assert max_rel(out, ref) <= 2e-5
    # and against fp32 torch on the bf16-rounded operands (the bf16 bar)
    h = (A.float() @ W1.float().t() + b1).to(torch.bfloat16).float()
    z = F.gelu(h).to(torch.bfloat16).float() @ W2.float().t() + b2
    s = dp.repeat_interleave(rps)[:M, None] if with_dp else 1.0
    assert max_rel(out, sc + s * (gamma * z)) <= 2e-2


@pytest.mark.parametrize("M,C", [(300, 96), (1000, 40), (128, 768), (5, 8)])
def test_split_operands_reconstruct_fp32(M, C):
    """cnx_split3 / weight-prep mode 3: [hi | mid | hi] and [hi | hi | mid]; hi + mid = x to 2^-16 relative."""
    from imageclassification_b200 import _lib as L
    lib = L.load()
    g = torch.Generator(device=DEV).manual_seed(M + C)
    x = torch.randn(M, C, device=DEV, generator=g) * 3
    a3 = torch.empty(M, 3 * C, dtype=torch.bfloat16, device=DEV)
    L.check(lib.cnx_split3(L.ptr(x), M, C, L.ptr(a3), 3, L.stream()), "split3")
    hi, mid, hi2 = a3[:, :C].float(), a3[:, C:2 * C].float(), a3[:, 2 * C:].float()
    assert torch.equal(hi, hi2) and torch.equal(hi, x.to(torch.bfloat16).float())
    a2 = torch.empty(M, 2 * C, dtype=torch.bfloat16, device=DEV)
    L.check(lib.cnx_split3(L.ptr(x), M, C, L.ptr(a2), 2, L.stream()), "split3(2 segments)")
    assert torch.equal(a2, a3[:, :2 * C])
    assert ((hi + mid - x).abs() <= x.abs() * 2 ** -16 + 1e-30).all()
    b3 = torch.empty(M, 3 * C, dtype=torch.bfloat16, device=DEV)
    L.check(lib.cnx_weight_prep(L.ptr(x), M, C, None, 3, L.ptr(b3), L.dt(torch.bfloat16), L.stream()), "weight_prep")
    assert torch.equal(b3[:, :C], a3[:, :C]) and torch.equal(b3[:, C:2 * C], a3[:, :C]) and torch.equal(b3[:, 2 * C:], a3[:, C:2 * C])


@pytest.mark.parametrize("M,C", [(3136, 96), (777, 128), (1000, 384), (130, 768)])
def test_x3_mlp_forward_is_fp32_accurate(M, C):
    """fc1 + GELU + fc2 + layer-scale + residual with split operands vs float64 on the SAME fp32 inputs: 1e-5 typical."""
    from imageclassification_b200 import _lib as L
    lib = L.load()
    A, W1, b1, W2, b2, gamma = _mk(M, C, torch.float32, 11 * M + C)
    sc = torch.randn(M, C, device=DEV)
    bf = torch.bfloat16
    a3 = torch.empty(M, 3 * C, dtype=bf, device=DEV)
    w13 = torch.empty(4 * C, 3 * C, dtype=bf, device=DEV)
    w23 = torch.empty(C, 12 * C, dtype=bf, device=DEV)
    g2 = torch.empty(M, 8 * C, dtype=bf, device=DEV)
    out = torch.empty(M, C, device=DEV)
    st = L.stream()
    L.check(lib.cnx_split3(L.ptr(A), M, C, L.ptr(a3), 3, st))
    L.check(lib.cnx_weight_prep(L.ptr(W1), 4 * C, C, None, 3, L.ptr(w13), L.dt(bf), st))
    L.check(lib.cnx_weight_prep(L.ptr(W2), C, 4 * C, None, 3, L.ptr(w23), L.dt(bf), st))
    L.check(lib.cnx_gemm_bias_gelu_fwd_x3(L.ptr(a3), L.ptr(w13), L.ptr(b1), M, 4 * C, 3 * C, L.ptr(g2), 3, st))
    # the two-segment A operand [hi | mid] with a wrapping K loop: identical MMAs, identical result
    a2 = a3[:, :2 * C].contiguous()
    g2w = torch.empty_like(g2)
    L.check(lib.cnx_gemm_bias_gelu_fwd_x3(L.ptr(a2), L.ptr(w13), L.ptr(b1), M, 4 * C, 3 * C, L.ptr(g2w), 2, st))
    assert torch.equal(g2w, g2)
    o3 = torch.empty(M, 4 * C, device=DEV)
    o2 = torch.empty_like(o3)
    L.check(lib.cnx_gemm_plain(L.ptr(a3), L.ptr(w13), L.ptr(b1), L.ptr(o3), L.dt(torch.float32), M, 4 * C, 3 * C, L.dt(bf), 0, st))
    L.check(lib.cnx_gemm_plain(L.ptr(a2), L.ptr(w13), L.ptr(b1), L.ptr(o2), L.dt(torch.float32), M, 4 * C, 3 * C, L.dt(bf),
                               L.CNX_GEMM_A_SPLIT2, st))
    assert torch.equal(o2, o3)
    assert max_rel(o3.double(), A.double() @ W1.double().t() + b1.double()) <= 2e-5
    gd = F.gelu(A.double() @ W1.double().t() + b1.double())
    ghat = g2[:, :4 * C].double() + g2[:, 4 * C:].double()
    assert max_rel(ghat, gd) <= 2e-5
    L.check(lib.cnx_gemm_bias_scale_residual_fwd(L.ptr(g2), L.ptr(w23), L.ptr(b2), L.ptr(gamma), None, 49, L.ptr(sc), L.ptr(out),
                                                 L.dt(torch.float32), M, C, 12 * C, L.dt(bf), L.CNX_GEMM_A_SPLIT2, st))
    ref = sc.double() + gamma.double() * (gd @ W2.double().t() + b2.double())
    assert max_rel(out.double() - sc.double(), ref - sc.double()) <= 3e-5
    # the same fc2 from an explicit three-segment operand (no wrap): identical MMAs, identical result
    g3 = torch.cat([g2, g2[:, :4 * C]], dim=1).contiguous()
    out3 = torch.empty_like(out)
    L.check(lib.cnx_gemm_bias_scale_residual_fwd(L.ptr(g3), L.ptr(w23), L.ptr(b2), L.ptr(gamma), None, 49, L.ptr(sc), L.ptr(out3),
                                                 L.dt(torch.float32), M, C, 12 * C, L.dt(bf), 0, st))
    assert torch.equal(out, out3)


@pytest.mark.parametrize("M,C", [(3136, 96), (1000, 128), (256, 192), (777, 192), (50176, 96)])
def test_dgrad_gelu_recompute_matches_saved_gelu_prime(M, C):
    """cnx_gemm_dgrad_gelu_recompute_bwd (GELU'(h) recomputed from xn, W1, b1 in a second TMEM accumulator) against the
    two-kernel path it replaces at the HBM-bound stages: fc1 saving GELU'(h), then cnx_gemm_dgrad_gelu_bwd.  Same MMAs, same
    roundings: bit-identical."""
    from cabi import gemm_bias_gelu, gemm_dgelu
    from imageclassification_b200 import _lib as L
    lib = L.load()
    bf = torch.bfloat16
    xn, W1, b1, W2, b2, gamma = _mk(M, C, bf, 5 * M + C)
    g = torch.Generator().manual_seed(M)
    dz = torch.randn(M, C, generator=g).to(bf).to(DEV)
    Bt = (gamma[:, None] * W2.float()).t().contiguous().to(bf)              # [4C, C] = (gamma . W2)^T
    gp, _ = gemm_bias_gelu(xn, W1, b1)
    ref = gemm_dgelu(dz, Bt, gp)
    dh = torch.empty_like(ref)
    L.check(lib.cnx_gemm_dgrad_gelu_recompute_bwd(L.ptr(dz), L.ptr(Bt), L.ptr(xn), L.ptr(W1), L.ptr(b1), L.ptr(dh), M, 4 * C, C,
                                                  L.dt(bf), L.stream()), "gemm_dgrad_gelu_recompute_bwd")
    assert torch.equal(dh, ref), max_rel(dh.float(), ref.float())
    # and against float64 math on the same bf16 inputs
    h = (xn.double() @ W1.double().t() + b1.double()).to(bf).double()
    gpd = 0.5 * (1 + torch.erf(h / math.sqrt(2))) + h * torch.exp(-0.5 * h * h) / math.sqrt(2 * math.pi)
    exact = (dz.double() @ Bt.double().t()) * gpd
    assert max_rel(dh.double(), exact) <= 2e-2


@pytest.mark.parametrize("M,C", [(3136, 96), (777, 128), (1000, 384), (130, 768), (50176, 96)])
def test_x3_training_epilogues_fc1_gelu_prime_and_dgelu_split(M, C):
    """fp32 training on split operands: cnx_gemm_bias_gelu_fwd_x3_train (g as [hi | mid] AND fp32 GELU'(h) from one epilogue) and
    cnx_gemm_dgrad_gelu_bwd_x3 ((dz . W2s) * GELU'(h) leaving as [hi | mid]) against float64 on the same fp32 inputs, and against
    the separate-pass kernels they replace (cnx_gemm_plain + cnx_gelu_split / cnx_mul_split)."""
    from imageclassification_b200 import _lib as L
    lib = L.load()
    A, W1, b1, W2, b2, gamma = _mk(M, C, torch.float32, 7 * M + C)
    bf, f32 = torch.bfloat16, torch.float32
    st = L.stream()
    N = 4 * C
    a2 = torch.empty(M, 2 * C, dtype=bf, device=DEV)
    w13 = torch.empty(N, 3 * C, dtype=bf, device=DEV)
    L.check(lib.cnx_split3(L.ptr(A), M, C, L.ptr(a2), 2, st))
    L.check(lib.cnx_weight_prep(L.ptr(W1), N, C, None, 3, L.ptr(w13), L.dt(bf), st))
    g2 = torch.empty(M, 2 * N, dtype=bf, device=DEV)
    gp = torch.empty(M, N, dtype=f32, device=DEV)
    L.check(lib.cnx_gemm_bias_gelu_fwd_x3_train(L.ptr(a2), L.ptr(w13), L.ptr(b1), M, N, 3 * C, L.ptr(g2), L.ptr(gp), 2, st))
    hd = (A.double() @ W1.double().t() + b1.double()).requires_grad_(True)
    gd = F.gelu(hd)
    gd.sum().backward()
    assert max_rel(g2[:, :N].double() + g2[:, N:].double(), gd.detach()) <= 2e-5
    assert max_rel(gp.double(), hd.grad) <= 4e-5          # the 1e-5 of the x3 pre-activation times max |GELU''| ~ 1.1 (19 M values)
    # the two-kernel path it replaces: same g pieces up to the erf implementation (A-S 7.1.26 vs erff), same GELU' to 1e-6
    h = torch.empty(M, N, dtype=f32, device=DEV)
    g2b = torch.empty_like(g2)
    L.check(lib.cnx_gemm_plain(L.ptr(a2), L.ptr(w13), L.ptr(b1), L.ptr(h), L.dt(f32), M, N, 3 * C, L.dt(bf), L.CNX_GEMM_A_SPLIT2, st))
    L.check(lib.cnx_gelu_split(L.ptr(h), M, N, L.ptr(g2b), st))
    assert max_rel(gp, h) <= 2e-6
    # (the pieces carry g to 2^-17: two g's that differ in the last fp32 bits can round their mid pieces differently)
    assert max_rel(g2[:, :N].float() + g2[:, N:].float(), g2b[:, :N].float() + g2b[:, N:].float()) <= 2e-5
    # backward: dh = (dz . (gamma W2)) * GELU'(h), split
    g = torch.Generator(device=DEV).manual_seed(M)
    dz = torch.randn(M, C, device=DEV, generator=g)
    dz2 = torch.empty(M, 2 * C, dtype=bf, device=DEV)
    L.check(lib.cnx_split3(L.ptr(dz), M, C, L.ptr(dz2), 2, st))
    w2s = (W2 * gamma[:, None]).t().contiguous()                     # [4C, C]
    w2s3 = torch.empty(N, 3 * C, dtype=bf, device=DEV)
    L.check(lib.cnx_weight_prep(L.ptr(w2s), N, C, None, 3, L.ptr(w2s3), L.dt(bf), st))
    dh2 = torch.empty(M, 2 * N, dtype=bf, device=DEV)
    L.check(lib.cnx_gemm_dgrad_gelu_bwd_x3(L.ptr(dz2), L.ptr(w2s3), L.ptr(gp), L.ptr(dh2), M, N, 3 * C, 2, st))
    ref = (dz.double() @ w2s.double().t()) * gp.double()
    assert max_rel(dh2[:, :N].double() + dh2[:, N:].double(), ref) <= 2e-5
    t1 = torch.empty(M, N, dtype=f32, device=DEV)
    dh2b = torch.empty_like(dh2)
    L.check(lib.cnx_gemm_plain(L.ptr(dz2), L.ptr(w2s3), None, L.ptr(t1), L.dt(f32), M, N, 3 * C, L.dt(bf), L.CNX_GEMM_A_SPLIT2, st))
    L.check(lib.cnx_mul_split(L.ptr(t1), L.ptr(gp), M, N, L.ptr(dh2b), st))
    assert torch.equal(dh2, dh2b)                                    # same MMAs, same fp32 multiply, same split


@pytest.mark.parametrize("M,N1,N2", [(3136, 96, 384), (1000, 384, 96), (50176, 384, 1536), (777, 768, 192), (4096, 1536, 384),
                                     (130, 3072, 768), (12544, 768, 3072), (5000, 200, 1000)])
@pytest.mark.parametrize("one_loop", [False, True], ids=["three_launches", "one_loop"])
def test_wgrad_x3_one_k_loop_over_three_products(M, N1, N2, one_loop):
    """cnx_gemm_wgrad_x3 / cnx_gemm_wgrad_x3_one_loop: out = X^T . Y and the column sums of X from split operands [hi | mid] — three
    launches (hi.hi, mid.hi, hi.mid), or ONE launch whose K loop walks the three products (single-CTA, CTA-pair 256 x 256 and
    256 x 384 tiles; ragged M, N1, N2) — against float64, with and without accumulation into an existing gradient.  The one-loop
    form has three times the accumulation-chain length in tensor memory: its bound is 8e-5 where the default's is 3e-5."""
    from imageclassification_b200 import _lib as L
    lib = L.load()
    fn = lib.cnx_gemm_wgrad_x3_one_loop if one_loop else lib.cnx_gemm_wgrad_x3
    tol = 8e-5 if one_loop else 3e-5
    g = torch.Generator(device=DEV).manual_seed(M + N1)
    X = torch.randn(M, N1, device=DEV, generator=g)
    Y = torch.randn(M, N2, device=DEV, generator=g)
    bf = torch.bfloat16
    st = L.stream()
    X2 = torch.empty(M, 2 * N1, dtype=bf, device=DEV)
    Y2 = torch.empty(M, 2 * N2, dtype=bf, device=DEV)
    L.check(lib.cnx_split3(L.ptr(X), M, N1, L.ptr(X2), 2, st))
    L.check(lib.cnx_split3(L.ptr(Y), M, N2, L.ptr(Y2), 2, st))
    wsb = lib.cnx_gemm_wgrad_workspace_bytes(M, N1, N2, L.dt(bf), 0)
    ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=DEV)
    out = torch.full((N1, N2), float("nan"), device=DEV)
    cs = torch.full((N1,), float("nan"), device=DEV)
    L.check(fn(L.ptr(X2), L.ptr(Y2), M, N1, N2, 0, L.ptr(out), L.ptr(cs), L.ptr(ws), wsb, st))
    ref = X.double().t() @ Y.double()
    assert max_rel(out.double(), ref) <= tol, max_rel(out.double(), ref)
    assert max_rel(cs.double(), X.double().sum(0)) <= 2e-5
    base = torch.randn(N1, N2, device=DEV, generator=g)
    acc = base.clone()
    cs2 = torch.ones(N1, device=DEV)
    L.check(fn(L.ptr(X2), L.ptr(Y2), M, N1, N2, 1, L.ptr(acc), L.ptr(cs2), L.ptr(ws), wsb, st))
    assert max_rel((acc - base).double(), ref) <= tol + 1e-5
    assert max_rel((cs2 - 1).double(), X.double().sum(0)) <= 3e-5
