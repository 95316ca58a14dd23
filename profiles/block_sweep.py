#!/usr/bin/env python
"""BASELINE config 5: isolated ConvNeXt Block sweep C in {96,192,384,768} x H=W in {56,28,14,7}, forward and forward+backward,
this package's Block (libcnx kernels) vs the same Block built from stock torch.nn modules (the oracle modules run on the GPU:
ATen / cuDNN / cuBLAS kernels — the "reference PyTorch GPU path").  bf16 autocast, fp32 residual stream, N chosen so that
N*H*W*C >= 64 Mi elements (>> L2).  CUDA events, 3 warm-ups, median of `--iters` runs.  One JSON line per shape.
usage: python profiles/block_sweep.py [--iters 5] [--mode bf16|fp32]"""
import argparse
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import imageclassification_b200 as P  # noqa: E402
from oracle import convnext as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32"])
ap.add_argument("--min-elems", type=int, default=64 << 20)
args = ap.parse_args()
dev = "cuda"
amp = args.mode == "bf16"


def time_block(mod, x, dout, backward):
    def run():
        if backward:
            xi = x.clone().requires_grad_(True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                y = mod(xi)
            y.backward(dout)
        else:
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                mod(x)
    for _ in range(3):
        run()
    for p in mod.parameters():
        p.grad = None
    torch.cuda.synchronize()
    ts = []
    for _ in range(args.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
        for p in mod.parameters():
            p.grad = None
    return statistics.median(ts)


for C in (96, 192, 384, 768):
    for H in (56, 28, 14, 7):
        N = max(2, -(-args.min_elems // (C * H * H)))
        N += N % 2
        torch.manual_seed(C + H)
        o = O.ConvNeXtBlock(C, drop_path=0.0, ls_init_value=1.0).to(dev)
        p = P.ConvNeXtBlock(C, drop_path=0.0, ls_init_value=1.0).to(dev)
        p.load_state_dict(o.state_dict())
        x = torch.randn(N, C, H, H, device=dev).contiguous(memory_format=torch.channels_last)
        dout = torch.randn(N, C, H, H, device=dev).contiguous(memory_format=torch.channels_last)
        rec = {"C": C, "H": H, "N": N, "mode": args.mode, "MC_Mi": round(N * C * H * H / 2 ** 20, 1)}
        for tag, bwd in (("fwd", False), ("fwd_bwd", True)):
            t_ref = time_block(o, x, dout, bwd)
            t_our = time_block(p, x, dout, bwd)
            rec[f"{tag}_torch_ms"] = round(t_ref, 3)
            rec[f"{tag}_libcnx_ms"] = round(t_our, 3)
            rec[f"{tag}_speedup"] = round(t_ref / t_our, 2)
        print(json.dumps(rec), flush=True)
        del o, p, x, dout
        torch.cuda.empty_cache()
