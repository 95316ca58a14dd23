#!/usr/bin/env python
"""Isolated timing of every libcnx kernel at the ConvNeXt-T / batch-256 shapes (BASELINE config 2 and 5).
CUDA events on the launching stream, 3 warm-ups, L2 flushed (256 MiB memset) before every timed launch.
Prints one JSON line per kernel: ms, algorithmic GB/s and TFLOP/s, fractions of MEASURED_PEAKS.json (burst peaks:
kernels timed alone).   usage: python profiles/kbench.py [--n 256] [--only dwconv,gemm,...] [--iters 5]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bench  # noqa: E402
import cabi  # noqa: E402
from imageclassification_b200 import _lib as L  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=256)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--only", default="")
ap.add_argument("--stages", default="0,1,2,3")
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--call-log", default=None, help="write [C-ABI name, bench label] of EVERY libcnx call of this run, in launch order "
                                                 "(to match an ncu launch list of the same command: profiles/make_traffic.py)")
args = ap.parse_args()
dev = "cuda"
peaks = bench._peaks()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
STAGES = [(96, 56), (192, 28), (384, 14), (768, 7)]
only = set(args.only.split(",")) if args.only else None


CALLS = []


class _Log:
    """stands in for a KernelTimer between timed sections: records every call's name and arguments, no events"""
    names = None

    class _Rec(list):
        def append(self, r):
            CALLS.append((r[0], r[3]))
    records = _Rec()


def timeit(tag, fn):
    L.TIMER = _Log if args.call_log else None
    for _ in range(args.warmup):
        fn()
    torch.cuda.synchronize()
    recs = []
    for _ in range(args.iters):
        flush.zero_()
        L.TIMER = L.KernelTimer()
        fn()
        torch.cuda.synchronize()
        recs.append(L.TIMER.summary())
        for name, lst in recs[-1].items():
            pass
        CALLS.extend((r[0], r[3]) for r in L.TIMER.records)
        L.TIMER = _Log if args.call_log else None
    # sum over the C-ABI calls of one fn() invocation, min over iterations
    best = None
    for r in recs:
        ms = sum(t for lst in r.values() for t, _ in lst)
        by = fl = 0
        for name, lst in r.items():
            for _, a in lst:
                b, f, _ = bench.kernel_work(name, a)
                by += b or 0
                fl += f
        if best is None or ms < best[0]:
            best = (ms, by, fl)
    ms, by, fl = best
    out = {"kernel": tag, "ms": round(ms, 4), "GBps": round(by / ms / 1e6, 1), "hbm_frac": round(by / ms / 1e6 / peaks["hbm"], 3),
           "TFLOPs": round(fl / ms / 1e9, 1), "tensor_frac": round(fl / ms / 1e9 / peaks["tensor_burst"], 3)}
    print(json.dumps(out), flush=True)


bf, f32 = torch.bfloat16, torch.float32
if args.call_log:
    L.TIMER = _Log
for si in [int(s) for s in args.stages.split(",")]:
    C, H = STAGES[si]
    N = args.n
    M = N * H * H
    g = torch.Generator(device=dev).manual_seed(si)
    x = torch.randn(N, H, H, C, device=dev, generator=g)
    w = torch.randn(C, 1, 7, 7, device=dev, generator=g) * 0.1
    b = torch.randn(C, device=dev, generator=g)
    lw, lb = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    tag = f"C{C}_H{H}"
    if only is None or "dwconv" in only:
        timeit(f"dwconv7_ln_fwd {tag}", lambda: cabi.dwconv7_ln_fwd(x, w, b, lw, lb, 1e-6, bf))
        timeit(f"dwconv7_ln_fwd_x3 {tag}", lambda: cabi.dwconv7_ln_fwd_x3(x, w, b, lw, lb, 1e-6, 2))
        dy = torch.randn(M, C, device=dev, generator=g).to(bf)
        timeit(f"dwconv7_dgrad {tag}", lambda: cabi.dwconv7_dgrad(dy, w, x, (N, H, H, C), f32))
        timeit(f"dwconv7_wgrad {tag}", lambda: cabi.dwconv7_wgrad(dy, x, P=max(1, L.load().cnx_sm_count() // (C // 32))))
        del dy
    if only is None or "lnf" in only:
        yb = torch.randn(M, C, device=dev, generator=g).to(bf)
        timeit(f"ln_fwd bf16->f32 {tag}", lambda: cabi.ln_fwd(yb, lw, lb, 1e-6, f32))
        xf = x.view(M, C)
        timeit(f"ln_fwd f32->bf16 {tag}", lambda: cabi.ln_fwd(xf, lw, lb, 1e-6, bf))
        del yb
    if only is None or "ln" in only:
        y = torch.randn(M, C, device=dev, generator=g).to(bf)
        dxn = torch.randn(M, C, device=dev, generator=g).to(bf)
        mean = torch.zeros(M, device=dev)
        rstd = torch.ones(M, device=dev)
        timeit(f"ln_bwd {tag}", lambda: cabi.ln_bwd(dxn, y, mean, rstd, lw, bf, P=int(os.environ.get("CNX_LN_P", "296"))))
        del y, dxn
    if only is None or "gemm" in only:
        A = torch.randn(M, C, device=dev, generator=g).to(bf)
        W1 = (torch.randn(4 * C, C, device=dev, generator=g) / C ** 0.5).to(bf)
        W2 = (torch.randn(C, 4 * C, device=dev, generator=g) / (4 * C) ** 0.5).to(bf)
        b1 = torch.zeros(4 * C, device=dev)
        b2 = torch.zeros(C, device=dev)
        gam = torch.ones(C, device=dev)
        hh, gg = cabi.gemm_bias_gelu(A, W1, b1)
        timeit(f"fc1_bias_gelu {tag}", lambda: cabi.gemm_bias_gelu(A, W1, b1))
        xs = x.view(M, C)
        if C <= 192:
            timeit(f"mlp_fused_fwd {tag}", lambda: cabi.mlp_fused_fwd(A, W1, b1, W2, b2, gam, None, H * H, xs))
        timeit(f"fc2_scale_res {tag}", lambda: cabi.gemm_scale_res(gg, W2, b2, gam, None, H * H, xs, f32))
        # fp32-accurate forward GEMMs on split operands (x3)
        lib = L.load()
        a3 = torch.randn(M, 3 * C, device=dev, generator=g).to(bf)
        w13 = (torch.randn(4 * C, 3 * C, device=dev, generator=g) / C ** 0.5).to(bf)
        w23 = (torch.randn(C, 12 * C, device=dev, generator=g) / (4 * C) ** 0.5).to(bf)
        g2 = torch.empty(M, 8 * C, dtype=bf, device=dev)
        o32 = torch.empty(M, C, device=dev)
        st = L.stream()
        timeit(f"fc1_gelu_x3 {tag}", lambda: L.check(lib.cnx_gemm_bias_gelu_fwd_x3(L.ptr(a3), L.ptr(w13), L.ptr(b1), M, 4 * C, 3 * C,
                                                                                     L.ptr(g2), 2, st)))
        timeit(f"fc2_scale_res_x3 {tag}", lambda: L.check(lib.cnx_gemm_bias_scale_residual_fwd(
            L.ptr(g2), L.ptr(w23), L.ptr(b2), L.ptr(gam), None, H * H, L.ptr(xs), L.ptr(o32), L.dt(f32), M, C, 12 * C, L.dt(bf),
            L.CNX_GEMM_A_SPLIT2, st)))
        if C == 96:
            a2 = a3[:, :2 * C].contiguous()
            timeit(f"mlp_fused_fwd_x3 {tag}", lambda: L.check(lib.cnx_mlp_fused_fwd_x3(
                L.ptr(a2), L.ptr(w13), L.ptr(b1), L.ptr(w23), L.ptr(b2), L.ptr(gam), None, H * H, L.ptr(xs), L.ptr(o32), M, C, st)))
            del a2
        del a3, g2, o32
        W2t = W2.t().contiguous()
        timeit(f"dgrad_fc2_gelu {tag}", lambda: cabi.gemm_dgelu(A, W2t, hh))
        W1t = W1.t().contiguous()
        timeit(f"dgrad_fc1 {tag}", lambda: cabi.gemm_plain(gg, W1t, None, bf))
        timeit(f"wgrad_fc2 {tag}", lambda: cabi.gemm_wgrad(A, gg))
        timeit(f"wgrad_fc1 {tag}", lambda: cabi.gemm_wgrad(gg, A))
        del A, hh, gg
    torch.cuda.empty_cache()

if args.call_log:
    L.TIMER = None
    with open(args.call_log, "w") as f:
        json.dump([[n, bench.kernel_work(n, a)[2]] for n, a in CALLS], f)
