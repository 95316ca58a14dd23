#!/usr/bin/env python
"""bench.py — BASELINE.json metric: ConvNeXt-T 224^2 bf16 training images/s (1/2/4/8 B200) + per-kernel roofline.

  python bench.py [--gpus N --steps K --warmup W]        this repo's CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference [...]                  the reference's CPU path (oracle port of engine.py's step) on
                                                           the host cores, bounded sample of the same workload

One "step" = one pass of the reference's training step (engine.py:27-97) over one batch of 256 synthetic images per
GPU: mixup(0.8)+label smoothing 0.1 -> ConvNeXt-T forward (bf16 autocast) -> SoftTargetCrossEntropy -> backward ->
AdamW -> ModelEmaV3 update -> the no-grad accuracy forward on the un-mixed batch (engine.py:89-97).  Prints ONE JSON
line (rank 0).  Keys: see the task contract; `value` = device-resident inputs, `e2e` = through engine.train_one_epoch
with pinned HOST buffers (H2D of the batch and D2H of the loss inside the timed region).
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ConvNeXt-T 224^2 bf16 train images/sec"
UNIT = "images/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="convnext_tiny")
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--classes", type=int, default=1000)
    ap.add_argument("--img", type=int, default=224)
    ap.add_argument("--cpu-sample-batch", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-breakdown", action="store_true")
    ap.add_argument("--torch-adamw", action="store_true", help="use torch.optim.AdamW instead of the fused AdamW+EMA")
    ap.add_argument("--no-prefetch", action="store_true", help="e2e leg: in-line H2D copies on the compute stream")
    ap.add_argument("--acc-autocast", action="store_true",
                    help="headline with the accuracy forward under bf16 autocast instead of fp32 outside autocast, where the "
                         "reference puts it (engine.py:89-97); by default that variant is only reported under `variants`")
    ap.add_argument("--no-amp", action="store_true",
                    help="fp32 everywhere (the reference's --use_amp false default) instead of the headline's bf16 autocast")
    ap.add_argument("--no-variants", action="store_true", help="skip the extra timed region of the `variants` key")
    ap.add_argument("--no-acc-forward", action="store_true", help="skip the reference's accuracy forward (not the default)")
    ap.add_argument("--update-freq", type=int, default=1,
                    help="gradient-accumulation micro-steps per optimizer step (the reference's --update_freq, engine.py:29,63-75); "
                         "one bench step = one optimizer step = this many micro-batches of --batch images per GPU")
    ap.add_argument("--cutmix", type=float, default=0.0, help="cutmix alpha (train.py:72-76; BASELINE config 3 uses 1.0)")
    ap.add_argument("--kernels-out", default=None,
                    help="side file for the per-shape kernel table / family table / launch timeline "
                         "(default gpurun_out/bench_kernels_N<gpus>.json)")
    return ap.parse_args()


def workload_name(a):
    uf = getattr(a, "update_freq", 1)
    return (f"{a.model} {a.img}x{a.img} {'fp32 (no autocast)' if getattr(a, 'no_amp', False) else 'bf16 autocast'}, batch {a.batch}/GPU"
            f"{f' x update_freq {uf}' if uf > 1 else ''}, {a.classes} classes, mixup 0.8{f' + cutmix {a.cutmix:g}' if getattr(a, 'cutmix', 0) else ''} + smoothing 0.1 + "
            f"SoftTargetCE + AdamW + ModelEmaV3(0.9995) + accuracy forward "
            f"{'under the same bf16 autocast' if getattr(a, 'acc_autocast', False) else 'in fp32 outside autocast as the reference places it'}"
            f" (engine.py:27-97)")


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed regions.  In-process NVML (nvidia_ml_py) every 50 ms: a
    query costs microseconds.  (Spawning `nvidia-smi` instead re-initialises NVML over every GPU of the box each time and was
    measured to stall this process's CUDA calls for ~100 ms — 10 ms per step over a 10-step region; kept only as a fallback.)"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    MASKS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index=0):
        self.rows, self.stop, self.index = [], threading.Event(), index
        self.primed = threading.Event()
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                if not uuid.startswith("GPU-"):
                    uuid = "GPU-" + uuid
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
                phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _sample_nvml(self):
        n, h = self.nvml, self.handle
        sm = float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM))
        try:
            pw = n.nvmlDeviceGetPowerUsage(h) / 1000.0
        except Exception:
            pw = 0.0
        try:
            mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(h))
        except Exception:
            mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(h))
        self.rows.append([str(sm), str(self.max_sm), str(pw)] + ["Active" if mask & m else "Not Active" for _, m in self.MASKS])

    def _run(self):
        while not self.stop.is_set():
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                    self.primed.set()
                else:
                    out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([c.strip() for c in out.split(",")])
                    self.primed.set()
            except Exception:
                pass
            self.stop.wait(float(os.environ.get("CNX_CLOCK_PERIOD", "0.05")) if self.nvml is not None else 1.0)

    def __enter__(self):
        # the FIRST NVML queries of a process are slow (tens of ms) and were seen to stall this process's kernel launches for one
        # step: let the sampler take its first sample before the timed region starts
        self.t.start()
        self.primed.wait(timeout=3.0)
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml / nvidia-smi unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        names = [n for n, _ in self.MASKS]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(self.rows[0][1]),
                "sm_mhz_min": min(sm) if sm else None,
                "power_w_max": max(float(r[2]) for r in self.rows), "samples": len(self.rows), "reasons": reasons,
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------------------ roofline
def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm": p["hbm_gbs"], "tensor_burst": p["bf16_tflops"], "tensor": p["bf16_tflops_sustained"], "src": "measured"}
    except Exception:
        return {"hbm": 6650.0, "tensor_burst": 1590.0, "tensor": 1400.0, "src": "fallback"}


def kernel_work(name, a):
    """(algorithmic bytes, flops, label) of one C-ABI call from its arguments (include/cnx.h order); DESIGN.md §roofline."""
    es = lambda d: 2 if d == 1 else 4
    if name == "cnx_dwconv7_ln_fwd":
        N, H, W, C = a[7:11]
        MC = N * H * W * C
        return MC * (es(a[1]) + 2 * es(a[13])) + 8 * N * H * W + 52 * C * 4, 106 * MC, f"dwconv7_ln_fwd C{C} H{H}"
    if name == "cnx_dwconv7_dgrad":
        N, H, W, C = a[6:10]
        MC = N * H * W * C
        return MC * (es(a[1]) + (2 if a[3] else 1) * es(a[5])), 98 * MC, f"dwconv7_dgrad C{C} H{H}"
    if name == "cnx_dwconv7_dgrad_dz":      # dy bf16 in, dres in, dx out (stream dtype) + the upstream Block's bf16 operand copy
        N, H, W, C = a[5:9]
        MC = N * H * W * C
        return MC * (2 + (2 if a[2] else 1) * es(a[4]) + 2), 98 * MC, f"dwconv7_dgrad C{C} H{H} +dz"
    if name == "cnx_dwconv7_wgrad":
        N, H, W, C = a[4:8]
        MC = N * H * W * C
        return MC * (es(a[1]) + es(a[3])) + 50 * C * 4, 98 * MC, f"dwconv7_wgrad C{C} H{H}"
    if name == "cnx_ln_bwd":
        M, C = a[7:9]
        return M * C * (es(a[1]) + es(a[3]) + es(a[10])) + 8 * M + 16 * C, 0, f"ln_bwd C{C}"
    if name == "cnx_ln_fwd_patch2":
        N, H, W, C = a[5:9]
        M = N * H * W
        return M * C * (es(a[1]) + es(a[10])) + 8 * M + 8 * C, 0, f"ln_fwd_patch2 C{C}"
    if name == "cnx_ln_bwd_patch2":
        N, H, W, C = a[7:11]
        M = N * H * W
        return M * C * (es(a[1]) + es(a[3]) + es(a[12])) + 8 * M + 16 * C, 0, f"ln_bwd_patch2 C{C}"
    if name == "cnx_patchify4_nchw":
        N, Cin, H, W = a[1:5]
        return N * Cin * H * W * (4 + es(a[6])), 0, "patchify4"
    if name == "cnx_ln_fwd":
        M, C = a[5:7]
        return M * C * (es(a[1]) + es(a[8])) + 8 * M + 8 * C, 0, f"ln_fwd C{C}"
    if name == "cnx_gemm_bias_gelu_fwd":
        M, N, K = a[3:6]
        e = es(a[8])
        return (M * K + N * K + M * N * (2 if a[6] else 1)) * e, 2 * M * N * K, f"fc1_gelu K{K}"
    if name == "cnx_gemm_bias_gelu_fwd_x3":
        # fp32-accurate GEMM on split bf16 operands: ALGORITHMIC flops are those of the fp32 product (2MNK); the kernel
        # executes 3x that on the tensor pipe by design (include/cnx.h "x3")
        M, N, K3 = a[3:6]
        return (M * (K3 // 3) * a[7] + N * K3 + M * 2 * N) * 2, 2 * M * N * (K3 // 3), f"fc1_gelu_x3 K{K3 // 3}"
    if name == "cnx_gemm_bias_gelu_fwd_x3_train":     # the same + the fp32 GELU'(h) tensor written by the epilogue
        M, N, K3 = a[3:6]
        return (M * (K3 // 3) * a[8] + N * K3 + M * 2 * N) * 2 + M * N * 4, 2 * M * N * (K3 // 3), f"fc1_gelu_x3_train K{K3 // 3}"
    if name == "cnx_gemm_dgrad_gelu_bwd_x3":          # dz pieces + split weights in, fp32 GELU' in, [hi | mid] of dh out
        M, N, K3 = a[4:7]
        return (M * (K3 // 3) * a[7] + N * K3 + M * 2 * N) * 2 + M * N * 4, 2 * M * N * (K3 // 3), f"dgrad_fc2_gelu_x3 K{K3 // 3}"
    if name == "cnx_dwconv7_ln_fwd_x3":
        N, H, W, C = a[6:10]
        MC = N * H * W * C
        return MC * (4 + 4 + 2 * a[14]) + 8 * N * H * W + 52 * C * 4, 106 * MC, f"dwconv7_ln_fwd_x3 C{C} H{H}"
    if name == "cnx_split3":
        M, C = a[1:3]
        return M * C * (4 + 2 * a[4]), 0, f"split3 C{C}"
    if name == "cnx_avgpool_nhwc_fwd":
        N, HW, C = a[2:5]
        return N * HW * C * es(a[1]) + N * C * 4, 0, f"avgpool_fwd C{C}"
    if name == "cnx_avgpool_nhwc_bwd":
        N, HW, C = a[1:4]
        return N * HW * C * es(a[5]) + N * C * 4, 0, f"avgpool_bwd C{C}"
    if name == "cnx_mixup_batch":
        B, C, H, W = a[2:6]
        return B * C * H * W * 4 * (3 if a[1] else 2), 0, "mixup_batch"
    if name == "cnx_gemm_bias_scale_residual_fwd":
        M, N, K = a[9:12]
        e, s = es(a[12]), es(a[8])
        if a[13] & 2:                                              # CNX_GEMM_A_SPLIT2 (x3): A holds 2 of the 3 K segments
            return (M * (K * 2 // 3) + N * K) * e + M * N * s * (2 if a[6] else 1), 2 * M * N * (K // 3), f"fc2_scale_res_x3 K{K // 3}"
        return (M * K + N * K) * e + M * N * s * (2 if a[6] else 1), 2 * M * N * K, f"fc2_scale_res K{K}"
    if name == "cnx_mlp_fused_fwd":
        M, C = a[10:12]
        return M * C * (2 + 4 + 4) + 8 * C * C * 2, 16 * M * C * C, f"mlp_fused_fwd C{C}"
    if name == "cnx_mlp_fused_fwd_x3":      # xn as [hi | mid] bf16 in, fp32 residual in, fp32 out; the split weights from L2
        M, C = a[10:12]
        return M * C * (4 + 4 + 4) + 2 * 4 * C * 3 * C * 2, 16 * M * C * C, f"mlp_fused_x3 C{C}"
    if name == "cnx_gemm_dgrad_gelu_bwd":
        M, N, K = a[4:7]
        e = es(a[7])
        return (M * K + N * K + 2 * M * N) * e, 2 * M * N * K, f"dgrad_fc2_gelu K{K}"
    if name == "cnx_gemm_dgrad_gelu_recompute_bwd":
        M, N, K = a[6:9]
        return (2 * M * K + 2 * N * K + M * N) * 2, 2 * M * N * K, f"dgrad_fc2_gelu_rc K{K}"     # algorithmic flops: the data gradient only
    if name == "cnx_gemm_plain":
        M, N, K = a[5:8]
        e = es(a[8])
        return (M * K + N * K) * e + M * N * es(a[4]), 2 * M * N * K, f"gemm_plain N{N} K{K}"
    if name == "cnx_gemm_wgrad":
        M, N1, N2 = a[2:5]
        e = es(a[10])
        return M * (N1 + N2) * e + N1 * N2 * 4, 2 * M * N1 * N2, f"wgrad {N1}x{N2}"
    if name in ("cnx_gemm_wgrad_x3", "cnx_gemm_wgrad_x3_one_loop"):
        M, N1, N2 = a[2:5]
        return 3 * M * (N1 + N2) * 2 + N1 * N2 * 4, 2 * M * N1 * N2, f"wgrad_x3 {N1}x{N2}"
    if name in ("cnx_gelu_split", "cnx_mul_split"):
        M, N = (a[1:3] if name == "cnx_gelu_split" else a[2:4])
        return M * N * (4 + 4 + 4), 0, f"{name[4:]} N{N}"
    if name == "cnx_grad_prep":
        M, C = a[4:6]
        return M * C * (es(a[1]) + es(a[7])), 0, f"grad_prep C{C}"
    if name in ("cnx_ema_lerp_multi",):
        return None, 0, "ema"
    return None, 0, name.replace("cnx_", "")


# ------------------------------------------------------------------------------------------------ ours
def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist

    import imageclassification_b200 as P
    from imageclassification_b200 import _lib as L, ddp as pddp, engine as pengine, optim as poptim

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl=ours) needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(88 + rank)
    np.random.seed(88 + rank)                                     # train.py:116-118

    model = P.create_model(a.model, pretrained=False, num_classes=a.classes, drop_path_rate=0.05).to(dev)
    n_params = sum(p.numel() for p in model.parameters())
    ema = P.ModelEmaV3(model, decay=0.9995, device=dev)           # train.py:201
    net = pddp.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
    groups = [{"params": [p for p in model.parameters() if p.requires_grad], "weight_decay": 5e-4}]   # optim_factory.py:23-47
    if a.torch_adamw:
        opt = torch.optim.AdamW(groups, lr=1e-3, weight_decay=0.0)
    else:
        opt = poptim.AdamW(groups, lr=1e-3, weight_decay=0.0)
        opt.fuse_ema(ema, model)
    mix = P.Mixup(mixup_alpha=0.8, cutmix_alpha=a.cutmix, label_smoothing=0.1, num_classes=a.classes)     # train.py:176-185
    crit = P.SoftTargetCrossEntropy()

    B = a.batch
    g = torch.Generator().manual_seed(1234 + rank)
    n_host = 4                                                    # distinct synthetic batches, cycled
    host = [(torch.randn(B, 3, a.img, a.img, generator=g).pin_memory(),
             torch.randint(0, a.classes, (B,), generator=g).pin_memory()) for _ in range(n_host)]
    devb = [(x.to(dev), t.to(dev)) for x, t in host]

    acc_fp32 = [not a.acc_autocast]

    def epoch(batches):
        if a.no_acc_forward:
            return _fast_epoch(batches)
        return pengine.train_one_epoch(net, crit, batches, opt, dev, 0, None, 0, ema, mix, update_freq=a.update_freq, use_amp=not a.no_amp,
                                       num_classes=a.classes, verbose=False, prefetch=not a.no_prefetch,
                                       acc_forward_fp32=acc_fp32[0], tune_gc=os.environ.get("CNX_ENGINE_TUNE_GC", "1") != "0")

    def _fast_epoch(batches):
        net.train(True)
        for x, t in batches:
            x, t = x.to(dev, non_blocking=True), t.to(dev, non_blocking=True)
            xs, ts = mix(x.clone(), t)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss = crit(net(xs), ts)
            lv = loss.item()
            loss.backward()
            opt.step()
            opt.zero_grad()
            ema.update(net)
        return {"loss": lv}

    class _Stamped(list):
        """the batch list, recording when the engine asks for each batch (host-side step intervals: a one-off stall of the
        launching thread shows up as one long interval, a slow link or slow kernels as uniformly long ones)"""

        def __iter__(self):
            self.t = []
            for b in list.__iter__(self):
                self.t.append(time.perf_counter())
                yield b

        def intervals_ms(self):
            return [1e3 * (b - a) for a, b in zip(self.t[:-1], self.t[1:])]

        def summary(self):
            iv = self.intervals_ms()
            if not iv:
                return None
            srt = sorted(iv)
            return {"median": round(srt[len(srt) // 2], 2), "max": round(srt[-1], 2), "argmax": iv.index(max(iv))}

    last_batches = [None]

    def timed(batches_src, steps):
        batches = _Stamped(batches_src[i % len(batches_src)] for i in range(steps * a.update_freq))
        last_batches[0] = batches
        # Python's cyclic garbage collector is switched off inside a timed region, as `timeit` does: a generation-2 pass was
        # measured to stop the launching thread for 40-190 ms at a fixed step of the run (one long host interval, always the
        # same index), which is an artefact of the process's object count, not of the step being measured
        gc.collect()
        if os.environ.get("CNX_BENCH_KEEP_GC", "0") != "1":         # (=1: leave the collector on, to measure the engine's own handling)
            gc.disable()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        stats = epoch(batches)
        e1.record()
        torch.cuda.synchronize()
        gc.enable()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, stats

    # host->device link check for the e2e leg: one standalone copy of a batch from the pinned buffers (GB/s), and whether the
    # buffers really are page-locked (a pageable source would make the "non-blocking" copy block the launching thread)
    torch.cuda.synchronize()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    scratch = torch.empty_like(devb[0][0])
    scratch.copy_(host[0][0], non_blocking=True)
    h0.record()
    scratch.copy_(host[1 % n_host][0], non_blocking=True)
    h1.record()
    torch.cuda.synchronize()
    h2d_gbps = round(host[0][0].numel() * 4 / (h0.elapsed_time(h1) * 1e-3) / 1e9, 1)
    pinned = all(x.is_pinned() and t.is_pinned() for x, t in host)
    del scratch
    if world > 1:                                                  # slowest link / any unpinned buffer over the ranks
        lt = torch.tensor([h2d_gbps, 1.0 if pinned else 0.0], device=dev, dtype=torch.float64)
        dist.all_reduce(lt, op=dist.ReduceOp.MIN)
        h2d_gbps, pinned = round(lt[0].item(), 1), bool(lt[1].item() > 0.5)

    # warm-up (both input sources), then the two timed regions
    t_warm = time.perf_counter()
    timed(devb, max(a.warmup, 3))
    timed(host, max(a.warmup, 3))
    # ... and keep the GPU under load until the board's power limiter has settled: about one second after sustained load starts
    # the limiter engages with a transient (one step of 40-190 ms at a fixed position of the run, whichever region was being
    # timed then); W warm-up steps of 35 ms end before it
    extra_warm = 0
    settle_s = float(os.environ.get("CNX_BENCH_SETTLE_S", "3.0"))

    def settled():
        el = time.perf_counter() - t_warm
        if world > 1:                                              # one decision for all ranks (timed() has a barrier)
            t = torch.tensor([el], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            el = t.item()
        return el >= settle_s

    while extra_warm < 120 and not settled():
        timed(devb, 4)
        extra_warm += 4                                 # (first host-fed steps grow the allocator's pools: seen as a 300 ms one-off)
    names_top = None
    with ClockSampler(local) as clk:
        c0 = L.gpu_launches()
        ms_dev, stats = timed(devb, a.steps)
        launches = L.gpu_launches() - c0
        dev_iv = last_batches[0].summary()
        ms_e2e, _ = timed(host, a.steps)
        host_iv = last_batches[0].summary()
        variants = None
        if not a.no_variants and not a.no_acc_forward and not a.no_amp:
            # the same step with the accuracy forward on the other side of the autocast boundary (see engine.py docstring)
            acc_fp32[0] = not acc_fp32[0]
            timed(devb, 3)
            ms_v, _ = timed(devb, a.steps)
            acc_fp32[0] = not acc_fp32[0]
            variants = {("accuracy_forward_fp32_outside_autocast" if a.acc_autocast else "accuracy_forward_under_bf16_autocast"):
                        {"value": round(B * world * a.steps * a.update_freq / (ms_v * 1e-3), 1), "unit": UNIT, "ms_per_step": round(ms_v / a.steps, 3)}}
    clocks = clk.summary()

    # per-kernel breakdown pass (untimed for the headline): every C-ABI call bracketed by CUDA events
    roof, table, families, timeline = None, [], [], None
    peaks = _peaks()
    if not a.no_breakdown:
        L.TIMER = L.KernelTimer()
        bsteps = 3
        L.TIMER.mark_base()
        ms_b, _ = timed(devb, bsteps)
        recs = L.TIMER.summary()
        tl = L.TIMER.timeline()
        L.TIMER = None
        table, families, roof = roofline_tables(recs, bsteps, peaks, ms_b)
        timeline = timeline_summary(tl, bsteps)

    imgs = B * world * a.steps * a.update_freq
    out = {
        "metric": METRIC, "value": round(imgs / (ms_dev * 1e-3), 1), "unit": UNIT, "n_gpus": world, "steps": a.steps,
        "warmup": max(a.warmup, 3), "ms_per_step": round(ms_dev / a.steps, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if a.no_amp else "bf16", "data": "synthetic (randn images, random labels; random-init weights)",
        "config": config_dict(a, world),
        "e2e": {"value": round(imgs / (ms_e2e * 1e-3), 1), "unit": UNIT,
                "h2d_bytes_per_step": (B * 3 * a.img * a.img * 4 + B * 8) * a.update_freq, "d2h_bytes_per_step": 4 * a.update_freq,
                "ms_per_step": round(ms_e2e / a.steps, 3), "host_buffers_pinned": pinned, "h2d_link_GBps": h2d_gbps,
                "api": "engine.train_one_epoch on pinned host batches"},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof, "families": families,
        "run": {"params": n_params, "final_loss": stats.get("loss"), "settle_warmup_steps": extra_warm,
                "optimizer": "torch.optim.AdamW" if a.torch_adamw else "libcnx fused AdamW+EMA",
                "residual_stream": "bf16 stages 1-3 (CNX_BF16_STREAM=1)" if os.environ.get("CNX_BF16_STREAM", "0") == "1" else "fp32 (reference promotion)",
                "host_step_interval_ms": {"resident": dev_iv, "e2e": host_iv}},
        "variants": variants,
    }
    if rank == 0:
        if not a.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_step_rate(a, steps=2, warmup=1)
        # the full per-shape kernel table, the per-family table and the launch timeline go to a side file: the JSON line stays
        # small enough for any log tail (round 1's 20 KB line was cut by the driver's tail and did not parse)
        side = a.kernels_out or os.path.join(ROOT, "gpurun_out", f"bench_kernels_N{world}.json")
        try:
            os.makedirs(os.path.dirname(side), exist_ok=True)
            with open(side, "w") as f:
                json.dump({"line": out, "kernels": table, "families_all": families_all(table), "timeline": timeline}, f, indent=1)
            out["kernels_file"] = os.path.relpath(side, ROOT)
        except OSError as e:
            out["kernels_file"] = f"not written: {e}"
        line = json.dumps(out, separators=(",", ":"))
        if len(line) > 3900:                                       # never let optional detail endanger the contract keys
            for k in ("variants", "run", "families"):
                out.pop(k, None)
                line = json.dumps(out, separators=(",", ":"))
                if len(line) <= 3900:
                    break
        print(line, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def config_dict(a, world):
    """`config` of the JSON line: identical for this arm and for `--impl reference` (the driver compares them)."""
    return {"workload": workload_name(a), "global_batch": a.batch * world * getattr(a, "update_freq", 1), "parallelism": f"dp{world}",
            "l2": "per-step activation footprint >> 126 MB L2 (inputs larger than L2, no explicit flush)"}


_NOT_TRAINING = ("_x3", "split3")          # the reference's fp32 accuracy forward (engine.py:89-97): in the step, not in fwd+bwd


def _family(label):
    return label.split(" ")[0]


def roofline_tables(recs, bsteps, peaks, ms_b):
    """per-shape table, per-family table and the `roofline` object (dominant training-path kernel family: all launches of
    one kernel entry point over all its shapes; achieved = total algorithmic bytes or flops / total CUDA-event time)."""
    agg = {}
    for name, lst in recs.items():
        for ms, args in lst:
            by, fl, label = kernel_work(name, args)
            d = agg.setdefault(label, {"ms": 0.0, "n": 0, "bytes": 0, "flops": 0, "name": name, "samples": []})
            d["samples"].append(ms)
            d["n"] += 1
            d["bytes"] += by or 0
            d["flops"] += fl
    for d in agg.values():
        # a launch that took more than 4x the median of its (kernel, shape) group is a one-off (seen: a single 2.8 ms launch of a
        # 51 us kernel inside a 3-step pass — a board-level hiccup, not the kernel): it counts as the median, and is reported
        med = statistics.median(d["samples"])
        d["outliers"] = sum(1 for v in d["samples"] if v > 4 * med)
        d["ms"] = sum(v if v <= 4 * med else med for v in d["samples"])
    tot = sum(d["ms"] for d in agg.values())
    table = []
    for label, d in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        gbs = d["bytes"] / (d["ms"] * 1e-3) / 1e9 if d["bytes"] else None
        tfs = d["flops"] / (d["ms"] * 1e-3) / 1e12 if d["flops"] else None
        is_gemm = d["name"].startswith("cnx_gemm") or d["name"].startswith("cnx_mlp")
        # the fp32-accurate split-operand GEMMs execute 3 MMAs per algorithmic product by design (include/cnx.h "x3"): their
        # tensor roof is set by the EXECUTED flops; the algorithmic fraction (SURVEY 8d flops) is reported beside it
        xf = 3 if (is_gemm and "_x3" in label) else 1
        # time the launch would take at the roof that bounds it (max of the HBM time and, for GEMMs, the tensor time)
        roof_ms = max(d["bytes"] / (peaks["hbm"] * 1e9), (xf * d["flops"] / (peaks["tensor"] * 1e12)) if is_gemm else 0.0) * 1e3
        table.append({"kernel": label, "family": _family(label), "calls_per_step": d["n"] / bsteps,
                      "ms_per_step": round(d["ms"] / bsteps, 4), "share_of_cnx": round(d["ms"] / tot, 4),
                      "bytes_per_launch": d["bytes"] // max(d["n"], 1), "flops_per_launch": d["flops"] // max(d["n"], 1),
                      "GBps": gbs and round(gbs, 1), "hbm_frac": gbs and round(gbs / peaks["hbm"], 4),
                      "TFLOPs": tfs and round(tfs, 2), "tensor_frac": (tfs and round(tfs / peaks["tensor"], 4)) if is_gemm else None,
                      "executed_flops_factor": xf, "outlier_launches": d["outliers"],
                      "roof_ms_per_step": round(roof_ms / bsteps, 4), "gemm": is_gemm})
    fams = families_all(table)
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tr = json.load(f)
    except Exception:
        tr = {"kernels": {}, "source": None}
    roof = None
    train = [f for f in fams if not any(t in f["family"] for t in _NOT_TRAINING) and f["frac"] is not None]
    if train:
        top = train[0]
        rows = [r for r in table if r["family"] == top["family"]]
        n = sum(r["calls_per_step"] for r in rows)
        ms = sum(r["ms_per_step"] for r in rows)
        if top["bound"] == "tensor":
            ach = sum(r["flops_per_launch"] * r["calls_per_step"] for r in rows) / (ms * 1e-3) / 1e12
            peak, unit = peaks["tensor"], "TFLOP/s"
        else:
            ach = sum(r["bytes_per_launch"] * r["calls_per_step"] for r in rows) / (ms * 1e-3) / 1e9
            peak, unit = peaks["hbm"], "GB/s"
        # DRAM bytes per launch (read + write) from the committed `ncu --set full` capture, launch-weighted over the family's shapes
        known = [r for r in rows if r["kernel"] in tr["kernels"]]
        traffic = None
        if known and sum(r["calls_per_step"] for r in known) >= 0.5 * n:
            traffic = round(sum(tr["kernels"][r["kernel"]] * r["calls_per_step"] for r in known) / sum(r["calls_per_step"] for r in known))
        roof = {"bound": top["bound"], "achieved": round(ach, 1), "peak": peak, "unit": unit, "frac": round(ach / peak, 4),
                "traffic": traffic, "kernel": top["family"], "launches_per_step": n, "avg_launch_ms": round(ms / n, 4),
                "ms_per_step": round(ms, 3), "share_of_step_kernels": top["share"], "roof_time_frac": top["roof_time_frac"],
                "algorithmic_bytes_per_launch": round(sum(r["bytes_per_launch"] * r["calls_per_step"] for r in rows) / n),
                "traffic_source": tr.get("source") if traffic else None,
                "peak_source": peaks["src"] + " MEASURED_PEAKS.json, sustained (timed inside the step)",
                "how": "all launches of the family over its shapes: sum of algorithmic work / sum of CUDA-event time",
                "cnx_kernels_ms_per_step": round(tot / bsteps, 3), "step_ms_with_events": round(ms_b / bsteps, 3)}
    compact = [{k: f[k] for k in ("family", "ms", "share", "bound", "frac")} for f in fams[:12]]
    return table, compact, roof


def families_all(table):
    """Aggregate the per-shape rows by kernel family.  frac = achieved / peak on the family's dominant roof (the one with the
    larger share of roof time); roof_time_frac = sum over shapes of the time at the bounding roof / measured time."""
    fam = {}
    tot = sum(r["ms_per_step"] for r in table) or 1.0
    for r in table:
        f = fam.setdefault(r["family"], {"ms": 0.0, "n": 0.0, "bytes": 0.0, "flops": 0.0, "xflops": 0.0, "roof_ms": 0.0,
                                         "gemm": r["gemm"]})
        f["ms"] += r["ms_per_step"]
        f["n"] += r["calls_per_step"]
        f["bytes"] += r["bytes_per_launch"] * r["calls_per_step"]
        f["flops"] += r["flops_per_launch"] * r["calls_per_step"]
        f["xflops"] += r["flops_per_launch"] * r["calls_per_step"] * r.get("executed_flops_factor", 1)
        f["roof_ms"] += r["roof_ms_per_step"]
    peaks = _peaks()
    out = []
    for name, f in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
        hb = f["bytes"] / (f["ms"] * 1e-3) / 1e9 / peaks["hbm"] if f["bytes"] and f["ms"] else None
        tf = f["flops"] / (f["ms"] * 1e-3) / 1e12 / peaks["tensor"] if (f["flops"] and f["gemm"] and f["ms"]) else None
        tfx = f["xflops"] / (f["ms"] * 1e-3) / 1e12 / peaks["tensor"] if (f["xflops"] and f["gemm"] and f["ms"]) else None
        if hb is None and tfx is None:
            bound, frac = None, None
        elif tfx is not None and tfx > (hb or 0):
            bound, frac = "tensor", round(tfx, 4)        # executed flops (= algorithmic except for the x3 split-operand GEMMs)
        else:
            bound, frac = "hbm", round(hb, 4)
        out.append({"family": name, "ms": round(f["ms"], 3), "launches": f["n"], "share": round(f["ms"] / tot, 4), "bound": bound,
                    "frac": frac, "hbm_frac": hb and round(hb, 4), "tensor_frac": tf and round(tf, 4),
                    "tensor_frac_executed": tfx and round(tfx, 4),
                    "roof_time_frac": round(f["roof_ms"] / f["ms"], 4) if f["ms"] and f["roof_ms"] else None})
    return out


def timeline_summary(tl, bsteps):
    """From (name, host issue, gpu start, gpu end) per C-ABI call: how much of the pass the GPU spent inside libcnx kernels, in
    the gaps between them (ATen kernels + idle), and how many calls found the GPU waiting for the launching thread."""
    if not tl:
        return None
    starved = [r for r in tl if r[2] - r[1] < 0.03]                  # kernel started < 30 us after the host issued it
    busy = sum(r[3] - r[2] for r in tl)
    gaps = [(tl[i][2] - tl[i - 1][3], tl[i][0], tl[i - 1][0]) for i in range(1, len(tl))]
    span = tl[-1][3] - tl[0][2]
    big = sorted(gaps, key=lambda g: -g[0])[:8]
    lead = [r[2] - r[1] for r in tl]
    by_name = {}
    for r in starved:
        by_name[r[0]] = by_name.get(r[0], 0) + 1
    return {"calls": len(tl), "span_ms_per_step": round(span / bsteps, 3), "cnx_busy_ms_per_step": round(busy / bsteps, 3),
            "gap_ms_per_step": round(sum(g[0] for g in gaps) / bsteps, 3),
            "starved_calls_per_step": round(len(starved) / bsteps, 1), "starved_by_call": by_name,
            "queue_lead_ms": {"median": round(sorted(lead)[len(lead) // 2], 3), "min": round(min(lead), 3), "max": round(max(lead), 3)},
            "largest_gaps_ms": [{"gap": round(g[0], 3), "before": g[1], "after": g[2]} for g in big],
            "rows": [[r[0], round(r[1], 3), round(r[2], 3), round(r[3], 3)] for r in tl[:len(tl) // bsteps]]}


# ------------------------------------------------------------------------------------------------ reference / CPU baseline
def cpu_step_rate(a, steps, warmup):
    """The oracle port of the reference step (oracle/engine.py <- engine.py:27-97) on the host cores, fp32, on a bounded
    sample (batch `--cpu-sample-batch`) of the same workload."""
    import numpy as np
    import torch

    from oracle import convnext as OC, ema as OE, engine as OEng, loss as OL, mixup as OM

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(88)
    np.random.seed(88)
    b = a.cpu_sample_batch
    model = OC.create_model(a.model, num_classes=a.classes, drop_path_rate=0.05)
    ema = OE.ModelEmaV3(model, decay=0.9995, device=torch.device("cpu"))
    opt = torch.optim.AdamW([{"params": list(model.parameters()), "weight_decay": 5e-4}], lr=1e-3, weight_decay=0.0)
    mix = OM.Mixup(mixup_alpha=0.8, label_smoothing=0.1, num_classes=a.classes)
    crit = OL.SoftTargetCrossEntropy()
    batch = [(torch.randn(b, 3, a.img, a.img), torch.randint(0, a.classes, (b,)))]
    for _ in range(warmup):
        OEng.train_one_epoch(model, crit, batch, opt, "cpu", 0, None, 0, ema, mix, num_classes=a.classes)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        OEng.train_one_epoch(model, crit, batch, opt, "cpu", 0, None, 0, ema, mix, num_classes=a.classes)
        ts.append(time.perf_counter() - t0)
    sec = sum(ts)
    return {"value": round(b * steps / sec, 2), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle/engine.py step, fp32, batch {b} (not {a.batch}) of the same workload, {steps} timed + {warmup} warm-up steps, "
                      f"{sec / steps:.2f} s/step", "ms_per_step": round(1e3 * sec / steps, 1)}


def run_reference(a):
    """The reference arm: the reference's own CPU implementation of the step (oracle port of engine.py:27-97; the reference
    itself cannot be installed: no setup.py, needs timm) on all host cores, EXACTLY --steps timed after --warmup untimed steps,
    each step a bounded sample (batch --cpu-sample-batch) of this arm's workload.  Same `config`, metric and unit as ours."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cb = cpu_step_rate(a, steps=max(1, a.steps), warmup=max(0, a.warmup))
    out = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": max(1, a.steps),
           "warmup": max(0, a.warmup), "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic (randn images, random labels; random-init weights)",
           "config": config_dict(a, max(1, a.gpus)),
           "note": "reference train.py CPU path (fp32: CUDA autocast does not apply on CPU); not installable (timm absent, no "
                   "network) -> oracle port of engine.py's step on the host cores; img/s = sample batch / step time",
           "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out, separators=(",", ":")), flush=True)


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
