"""Thin test helpers that call libcnx.so through the C-ABI (include/cnx.h) on torch CUDA tensors."""
import torch

from imageclassification_b200 import _lib as L


def _st():
    return L.stream()


def tap_major(w):
    lib = L.load()
    C = w.shape[0]
    wt = torch.empty((49, C), dtype=torch.float32, device=w.device)
    L.check(lib.cnx_dwconv7_weight_prep(L.ptr(w.contiguous()), C, L.ptr(wt), _st()), "dwconv7_weight_prep")
    return wt


def dwconv7_ln_fwd(x_nhwc, w, b, ln_w, ln_b, eps, act_dtype):
    lib = L.load()
    w = tap_major(w)
    N, H, W, C = x_nhwc.shape
    M = N * H * W
    y = torch.empty((M, C), dtype=act_dtype, device=x_nhwc.device)
    xn = torch.empty_like(y)
    mean = torch.empty((M,), dtype=torch.float32, device=x_nhwc.device)
    rstd = torch.empty_like(mean)
    L.check(lib.cnx_dwconv7_ln_fwd(L.ptr(x_nhwc), L.dt(x_nhwc), L.ptr(w), L.ptr(b), L.ptr(ln_w), L.ptr(ln_b), eps, N, H, W, C,
                                   L.ptr(y), L.ptr(xn), L.dt(act_dtype), L.ptr(mean), L.ptr(rstd), _st()), "dwconv7_ln_fwd")
    return y, xn, mean, rstd


def dwconv7_ln_fwd_x3(x_nhwc, w, b, ln_w, ln_b, eps, segments=3):
    """fp32 stream -> split operand [hi | mid | hi] bf16 [M, 3C] (+ the fp32 conv output used as scratch, mean, rstd)"""
    lib = L.load()
    w = tap_major(w)
    N, H, W, C = x_nhwc.shape
    M = N * H * W
    y = torch.empty((M, C), dtype=torch.float32, device=x_nhwc.device)
    a3 = torch.empty((M, segments * C), dtype=torch.bfloat16, device=x_nhwc.device)
    mean = torch.empty((M,), dtype=torch.float32, device=x_nhwc.device)
    rstd = torch.empty_like(mean)
    L.check(lib.cnx_dwconv7_ln_fwd_x3(L.ptr(x_nhwc), L.ptr(w), L.ptr(b), L.ptr(ln_w), L.ptr(ln_b), eps, N, H, W, C, L.ptr(y),
                                      L.ptr(a3), L.ptr(mean), L.ptr(rstd), segments, _st()), "dwconv7_ln_fwd_x3")
    return y, a3, mean, rstd


def ln_fwd(x2, w, b, eps, out_dtype):
    lib = L.load()
    M, C = x2.shape
    out = torch.empty((M, C), dtype=out_dtype, device=x2.device)
    mean = torch.empty((M,), dtype=torch.float32, device=x2.device)
    rstd = torch.empty_like(mean)
    L.check(lib.cnx_ln_fwd(L.ptr(x2), L.dt(x2), L.ptr(w), L.ptr(b), eps, M, C, L.ptr(out), L.dt(out_dtype), L.ptr(mean),
                           L.ptr(rstd), _st()), "ln_fwd")
    return out, mean, rstd


def ln_bwd(dxn, y, mean, rstd, ln_w, dy_dtype, P=64):
    lib = L.load()
    M, C = y.shape
    dy = torch.empty((M, C), dtype=dy_dtype, device=y.device)
    part = torch.empty((P, 2 * C), dtype=torch.float32, device=y.device)
    L.check(lib.cnx_ln_bwd(L.ptr(dxn), L.dt(dxn), L.ptr(y), L.dt(y), L.ptr(mean), L.ptr(rstd), L.ptr(ln_w), M, C, L.ptr(dy),
                           L.dt(dy_dtype), L.ptr(part), P, _st()), "ln_bwd")
    dwb = torch.empty((2 * C,), dtype=torch.float32, device=y.device)
    L.check(lib.cnx_reduce_partials(L.ptr(part), P, 2 * C, 1.0, 0, L.ptr(dwb), _st()), "reduce_partials")
    return dy, dwb[:C], dwb[C:]


def dwconv7_dgrad(dy2, w, dres_nhwc, shape, stream_dtype):
    lib = L.load()
    w = tap_major(w)
    N, H, W, C = shape
    dx = torch.empty((N, H, W, C), dtype=stream_dtype, device=dy2.device)
    L.check(lib.cnx_dwconv7_dgrad(L.ptr(dy2), L.dt(dy2), L.ptr(w), L.ptr(dres_nhwc), L.ptr(dx), L.dt(stream_dtype), N, H, W, C,
                                  _st()), "dwconv7_dgrad")
    return dx


def dwconv7_dgrad_dz(dy2, w, dres_nhwc, shape, stream_dtype, dp_up):
    """dx and the bf16 operand copy dz_up = bf16(dp_up[n] * dx) of the upstream Block, in one launch."""
    lib = L.load()
    w = tap_major(w)
    N, H, W, C = shape
    dx = torch.empty((N, H, W, C), dtype=stream_dtype, device=dy2.device)
    dz = torch.empty((N * H * W, C), dtype=torch.bfloat16, device=dy2.device)
    L.check(lib.cnx_dwconv7_dgrad_dz(L.ptr(dy2), L.ptr(w), L.ptr(dres_nhwc), L.ptr(dx), L.dt(stream_dtype), N, H, W, C, L.ptr(dz),
                                     L.ptr(dp_up), _st()), "dwconv7_dgrad_dz")
    return dx, dz


def grad_prep(dout_nhwc, dp, act_dtype):
    lib = L.load()
    N, H, W, C = dout_nhwc.shape
    dz = torch.empty((N * H * W, C), dtype=act_dtype, device=dout_nhwc.device)
    L.check(lib.cnx_grad_prep(L.ptr(dout_nhwc), L.dt(dout_nhwc), L.ptr(dp), H * W, N * H * W, C, L.ptr(dz), L.dt(act_dtype), _st()),
            "grad_prep")
    return dz


def dwconv7_wgrad(dy2, x_nhwc, P=32):
    lib = L.load()
    N, H, W, C = x_nhwc.shape
    part = torch.empty((P, 50, C), dtype=torch.float32, device=dy2.device)
    L.check(lib.cnx_dwconv7_wgrad(L.ptr(dy2), L.dt(dy2), L.ptr(x_nhwc), L.dt(x_nhwc), N, H, W, C, L.ptr(part), P, _st()),
            "dwconv7_wgrad")
    dw = torch.empty((C, 1, 7, 7), dtype=torch.float32, device=dy2.device)
    db = torch.empty((C,), dtype=torch.float32, device=dy2.device)
    L.check(lib.cnx_dwconv7_wgrad_finalize(L.ptr(part), P, C, 0, L.ptr(dw), L.ptr(db), _st()), "dwconv7_wgrad_finalize")
    return dw, db


def gemm_bias_gelu(A, W1, b1, flags=0):
    lib = L.load()
    M, K = A.shape
    N = W1.shape[0]
    h = torch.empty((M, N), dtype=A.dtype, device=A.device)
    g = torch.empty_like(h)
    L.check(lib.cnx_gemm_bias_gelu_fwd(L.ptr(A), L.ptr(W1), L.ptr(b1), M, N, K, L.ptr(h), L.ptr(g), L.dt(A), flags, _st()),
            "gemm_bias_gelu_fwd")
    return h, g


def gemm_scale_res(A, W2, b2, gamma, dp, rps, shortcut, stream_dtype, flags=0):
    lib = L.load()
    M, K = A.shape
    N = W2.shape[0]
    out = torch.empty((M, N), dtype=stream_dtype, device=A.device)
    L.check(lib.cnx_gemm_bias_scale_residual_fwd(L.ptr(A), L.ptr(W2), L.ptr(b2), L.ptr(gamma), L.ptr(dp), rps, L.ptr(shortcut),
                                                 L.ptr(out), L.dt(stream_dtype), M, N, K, L.dt(A), flags, _st()),
            "gemm_bias_scale_residual_fwd")
    return out


def gemm_dgelu(dz, Bt, h, flags=0):
    lib = L.load()
    M, K = dz.shape
    N = Bt.shape[0]
    dh = torch.empty((M, N), dtype=dz.dtype, device=dz.device)
    L.check(lib.cnx_gemm_dgrad_gelu_bwd(L.ptr(dz), L.ptr(Bt), L.ptr(h), L.ptr(dh), M, N, K, L.dt(dz), flags, _st()),
            "gemm_dgrad_gelu_bwd")
    return dh


def gemm_plain(A, B, bias, out_dtype, flags=0):
    lib = L.load()
    M, K = A.shape
    N = B.shape[0]
    out = torch.empty((M, N), dtype=out_dtype, device=A.device)
    L.check(lib.cnx_gemm_plain(L.ptr(A), L.ptr(B), L.ptr(bias), L.ptr(out), L.dt(out_dtype), M, N, K, L.dt(A), flags, _st()),
            "gemm_plain")
    return out


def gemm_wgrad(X, Y, flags=0, colsum=True):
    lib = L.load()
    M, N1 = X.shape
    N2 = Y.shape[1]
    d = L.dt(X)
    wsb = lib.cnx_gemm_wgrad_workspace_bytes(M, N1, N2, d, flags)
    ws = torch.empty(max(wsb // 4, 1), dtype=torch.float32, device=X.device)
    out = torch.empty((N1, N2), dtype=torch.float32, device=X.device)
    cs = torch.empty((N1,), dtype=torch.float32, device=X.device) if colsum else None
    L.check(lib.cnx_gemm_wgrad(L.ptr(X), L.ptr(Y), M, N1, N2, 0, L.ptr(out), L.ptr(cs), L.ptr(ws), wsb, d, flags, _st()),
            "gemm_wgrad")
    return out, cs


def rel_err(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def max_rel(a, b):
    """max |a-b| / max |b| — the 'relative on logits and gradients' measure of BASELINE.json north_star."""
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def mlp_fused_fwd(xn, W1, b1, W2, b2, gamma, dp, rps, shortcut):
    lib = L.load()
    M, C = xn.shape
    out = torch.empty((M, C), dtype=torch.float32, device=xn.device)
    L.check(lib.cnx_mlp_fused_fwd(L.ptr(xn), L.ptr(W1), L.ptr(b1), L.ptr(W2), L.ptr(b2), L.ptr(gamma), L.ptr(dp), rps,
                                  L.ptr(shortcut), L.ptr(out), M, C, _st()), "mlp_fused_fwd")
    return out
