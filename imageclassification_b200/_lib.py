"""ctypes binding of libcnx.so (include/cnx.h).

This is the only place Python touches the C-ABI.  There is no CPU fallback: if the library is missing
or a call fails, a RuntimeError is raised (the reference's engine.train_one_epoch would otherwise keep
training on silently wrong numbers).
"""
from __future__ import annotations

import ctypes
import os
import time
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int64, c_uint, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CNX_LIB", os.path.join(_HERE, "lib", "libcnx.so"))   # CNX_LIB: kernel-variant experiments

CNX_F32, CNX_BF16 = 0, 1
CNX_GEMM_FORCE_SIMT = 1
CNX_GEMM_A_SPLIT2 = 2
CNX_GEMM_OUT_ROUND_BF16 = 4
CNX_EMA_CHUNK = 8192

_lib = None


class EmaEntry(ctypes.Structure):
    _fields_ = [("ema", c_void_p), ("param", c_void_p), ("numel", c_int64), ("chunk_start", c_int64)]


class WeightPrepEntry(ctypes.Structure):
    _fields_ = [("W", c_void_p), ("row_scale", c_void_p), ("out", c_void_p), ("R", c_int64), ("Cc", c_int64),
                ("mode", ctypes.c_int32), ("out_dtype", ctypes.c_int32), ("tile_start", c_int64), ("tiles_x", c_int64)]


class AdamWEntry(ctypes.Structure):
    _fields_ = [("param", c_void_p), ("grad", c_void_p), ("exp_avg", c_void_p), ("exp_avg_sq", c_void_p),
                ("ema", c_void_p), ("numel", c_int64), ("chunk_start", c_int64)]


# name -> (restype, argtypes); must list every symbol include/cnx.h declares (tests/test_cabi.py checks)
_P, _I, _L, _F, _D = c_void_p, c_int, c_int64, c_float, c_double
SIGNATURES = {
    "cnx_version": (c_int, []),
    "cnx_last_error_string": (c_char_p, []),
    "cnx_sm_count": (c_int, []),
    "cnx_ema_lerp_multi": (c_int, [_P, _I, _L, _F, _P]),
    "cnx_adamw_ema_multi": (c_int, [_P, _I, _L, _D, _D, _D, _D, _D, _D, _D, _F, _P]),
    "cnx_grad_sumsq_multi": (c_int, [_P, _I, _L, _P, _F, _P, _P]),
    "cnx_scale_multi": (c_int, [_P, _I, _L, _P, _P]),
    "cnx_soft_target_ce_fwd": (c_int, [_P, _I, _P, _L, _L, _P, _P, _P, _P, _P]),
    "cnx_soft_target_ce_bwd": (c_int, [_P, _I, _P, _P, _P, _L, _L, _P, _I, _P]),
    "cnx_mixup_target": (c_int, [_P, _L, _L, _D, _D, _P, _P]),
    "cnx_mixup_batch": (c_int, [_P, _P, _L, _L, _L, _L, _D, _I, _I, _I, _I, _I, _P]),
    "cnx_dwconv7_weight_prep": (c_int, [_P, _L, _P, _P]),
    "cnx_dwconv7_ln_fwd": (c_int, [_P, _I, _P, _P, _P, _P, _F, _L, _L, _L, _L, _P, _P, _I, _P, _P, _P]),
    "cnx_ln_fwd": (c_int, [_P, _I, _P, _P, _F, _L, _L, _P, _I, _P, _P, _P]),
    "cnx_ln_bwd": (c_int, [_P, _I, _P, _I, _P, _P, _P, _L, _L, _P, _I, _P, _I, _P]),
    "cnx_reduce_partials": (c_int, [_P, _I, _L, _F, _I, _P, _P]),
    "cnx_reduce_partials_split": (c_int, [_P, _I, _L, _L, _I, _P, _P, _P]),
    "cnx_dwconv7_dgrad": (c_int, [_P, _I, _P, _P, _P, _I, _L, _L, _L, _L, _P]),
    "cnx_dwconv7_dgrad_dz": (c_int, [_P, _P, _P, _P, _I, _L, _L, _L, _L, _P, _P, _P]),
    "cnx_dwconv7_wgrad": (c_int, [_P, _I, _P, _I, _L, _L, _L, _L, _P, _I, _P]),
    "cnx_dwconv7_wgrad_finalize": (c_int, [_P, _I, _L, _I, _P, _P, _P]),
    "cnx_gemm_bias_gelu_fwd": (c_int, [_P, _P, _P, _L, _L, _L, _P, _P, _I, _I, _P]),
    "cnx_gemm_bias_scale_residual_fwd": (c_int, [_P, _P, _P, _P, _P, _L, _P, _P, _I, _L, _L, _L, _I, _I, _P]),
    "cnx_split3": (c_int, [_P, _L, _L, _P, _I, _P]),
    "cnx_dwconv7_ln_fwd_x3": (c_int, [_P, _P, _P, _P, _P, _F, _L, _L, _L, _L, _P, _P, _P, _P, _I, _P]),
    "cnx_gemm_bias_gelu_fwd_x3": (c_int, [_P, _P, _P, _L, _L, _L, _P, _I, _P]),
    "cnx_gemm_bias_gelu_fwd_x3_train": (c_int, [_P, _P, _P, _L, _L, _L, _P, _P, _I, _P]),
    "cnx_gemm_dgrad_gelu_bwd_x3": (c_int, [_P, _P, _P, _P, _L, _L, _L, _I, _P]),
    "cnx_gelu_split": (c_int, [_P, _L, _L, _P, _P]),
    "cnx_mul_split": (c_int, [_P, _P, _L, _L, _P, _P]),
    "cnx_gemm_wgrad_x3": (c_int, [_P, _P, _L, _L, _L, _I, _P, _P, _P, _L, _P]),
    "cnx_gemm_wgrad_x3_one_loop": (c_int, [_P, _P, _L, _L, _L, _I, _P, _P, _P, _L, _P]),
    "cnx_mlp_fused_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _L, _P, _P, _L, _L, _P]),
    "cnx_mlp_fused_fwd_x3": (c_int, [_P, _P, _P, _P, _P, _P, _P, _L, _P, _P, _L, _L, _P]),
    "cnx_gemm_dgrad_gelu_bwd": (c_int, [_P, _P, _P, _P, _L, _L, _L, _I, _I, _P]),
    "cnx_gemm_dgrad_gelu_recompute_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _L, _L, _L, _I, _P]),
    "cnx_gemm_plain": (c_int, [_P, _P, _P, _P, _I, _L, _L, _L, _I, _I, _P]),
    "cnx_gemm_wgrad_workspace_bytes": (c_int64, [_L, _L, _L, _I, _I]),
    "cnx_gemm_wgrad": (c_int, [_P, _P, _L, _L, _L, _I, _P, _P, _P, _L, _I, _I, _P]),
    "cnx_grad_prep": (c_int, [_P, _I, _P, _L, _L, _L, _P, _I, _P]),
    "cnx_weight_prep": (c_int, [_P, _L, _L, _P, _I, _P, _I, _P]),
    "cnx_weight_prep_multi": (c_int, [_P, _I, _L, _P]),
    "cnx_layerscale_finalize": (c_int, [_P, _P, _P, _P, _P, _L, _L, _I, _P, _P, _P, _P]),
    "cnx_cast_f32_to_bf16": (c_int, [_P, _L, _P, _P]),
    "cnx_avgpool_nhwc_fwd": (c_int, [_P, _I, _L, _L, _L, _P, _P]),
    "cnx_avgpool_nhwc_bwd": (c_int, [_P, _L, _L, _L, _P, _I, _P]),
    "cnx_patchify4_nchw": (c_int, [_P, _L, _L, _L, _L, _P, _I, _P]),
    "cnx_patch2": (c_int, [_P, _I, _L, _L, _L, _L, _P, _I, _P]),
    "cnx_ln_fwd_patch2": (c_int, [_P, _I, _P, _P, _F, _L, _L, _L, _L, _P, _I, _P, _P, _P]),
    "cnx_ln_bwd_patch2": (c_int, [_P, _I, _P, _I, _P, _P, _P, _L, _L, _L, _L, _P, _I, _P, _I, _P]),
}


# C-ABI calls that launch kernels, and how many kernels one call launches (for bench.py's `gpu_launches`;
# cnx_gemm_wgrad launches the GEMM + one partial reduction, + one more when the bias gradient is requested).
KERNELS_PER_CALL = {
    "cnx_ema_lerp_multi": 1, "cnx_adamw_ema_multi": 1, "cnx_grad_sumsq_multi": 2, "cnx_scale_multi": 1, "cnx_soft_target_ce_fwd": 1, "cnx_soft_target_ce_bwd": 1,
    "cnx_mixup_target": 1, "cnx_mixup_batch": 1, "cnx_dwconv7_ln_fwd": 1, "cnx_ln_fwd": 1, "cnx_ln_bwd": 1, "cnx_reduce_partials": 1, "cnx_reduce_partials_split": 1,
    "cnx_dwconv7_dgrad": 1, "cnx_dwconv7_dgrad_dz": 1, "cnx_dwconv7_wgrad": 1, "cnx_dwconv7_wgrad_finalize": 1, "cnx_dwconv7_weight_prep": 1, "cnx_gemm_bias_gelu_fwd": 1,
    "cnx_gemm_bias_scale_residual_fwd": 1, "cnx_gemm_dgrad_gelu_bwd": 1, "cnx_gemm_dgrad_gelu_recompute_bwd": 1, "cnx_gemm_plain": 1, "cnx_gemm_wgrad": 2, "cnx_gemm_wgrad_x3": 6, "cnx_gemm_wgrad_x3_one_loop": 2, "cnx_gelu_split": 1, "cnx_mul_split": 1,
    "cnx_grad_prep": 1, "cnx_weight_prep": 1, "cnx_weight_prep_multi": 1, "cnx_mlp_fused_fwd": 1, "cnx_mlp_fused_fwd_x3": 1, "cnx_split3": 1, "cnx_dwconv7_ln_fwd_x3": 1, "cnx_gemm_bias_gelu_fwd_x3": 1, "cnx_gemm_bias_gelu_fwd_x3_train": 1, "cnx_gemm_dgrad_gelu_bwd_x3": 1, "cnx_layerscale_finalize": 1, "cnx_cast_f32_to_bf16": 1,
    "cnx_avgpool_nhwc_fwd": 1, "cnx_avgpool_nhwc_bwd": 1, "cnx_patchify4_nchw": 1, "cnx_patch2": 1, "cnx_ln_fwd_patch2": 1, "cnx_ln_bwd_patch2": 1,
}
CALL_COUNTS = {k: 0 for k in KERNELS_PER_CALL}


class KernelTimer:
    """Optional per-call CUDA-event timing of selected C-ABI entry points (bench.py's roofline leg).  Events are
    recorded on torch's current stream, which is the stream every launch is given."""

    def __init__(self, names=None):
        self.names = set(names) if names is not None else None
        self.records = []   # (name, start_event, end_event, int args, host issue time)
        self.base = None    # (event, host time) set by mark_base(): origin of timeline()

    def mark_base(self):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self.base = (ev, time.perf_counter())

    def summary(self):
        out = {}
        for name, s, e, a, _ in self.records:
            out.setdefault(name, []).append((s.elapsed_time(e), a))
        return out

    def timeline(self):
        """[(name, host issue ms, gpu start ms, gpu end ms)] relative to mark_base(): `gpu start - host issue` close to zero
        means the GPU was waiting for the launching thread at that call (host-bound), a large lead means it was queued."""
        ev0, t0 = self.base
        return [(name, 1e3 * (th - t0), ev0.elapsed_time(s), ev0.elapsed_time(e)) for name, s, e, _, th in self.records]


TIMER: KernelTimer | None = None


def _wrap(name, fn):
    def call(*a):
        CALL_COUNTS[name] += 1
        t = TIMER
        if t is not None and (t.names is None or name in t.names):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            th = time.perf_counter()
            s.record()
            rc = fn(*a)
            e.record()
            t.records.append((name, s, e, a, th))
            return rc
        return fn(*a)
    call.__name__ = name
    return call


def gpu_launches() -> int:
    return sum(CALL_COUNTS[k] * KERNELS_PER_CALL[k] for k in CALL_COUNTS)


class _Lib:
    pass


def load():
    """Load libcnx.so once; raise (never fall back) if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"libcnx.so not found at {LIB_PATH}: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C imageclassification_b200/csrc`). There is no CPU fallback.")
    cdll = ctypes.CDLL(LIB_PATH)
    lib = _Lib()
    lib.cdll = cdll
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(cdll, name)
        fn.restype = res
        fn.argtypes = args
        setattr(lib, name, _wrap(name, fn) if name in KERNELS_PER_CALL else fn)
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().cnx_last_error_string()
        raise RuntimeError(f"libcnx {what} failed (code {rc}): {msg.decode() if msg else ''}")


def dt(t_or_dtype) -> int:
    d = t_or_dtype.dtype if isinstance(t_or_dtype, torch.Tensor) else t_or_dtype
    if d == torch.float32:
        return CNX_F32
    if d == torch.bfloat16:
        return CNX_BF16
    raise TypeError(f"libcnx supports float32 and bfloat16 tensors, got {d}")


def ptr(t):
    return None if t is None else t.data_ptr()


def stream(device=None) -> int:
    """torch's current stream on `device` (default: the current device).  Ops pass their tensors' device and launch under
    `torch.cuda.device(...)` when it is not the current one."""
    return torch.cuda.current_stream(device).cuda_stream


def require_same_device(*tensors) -> None:
    """Kernels are launched on the CURRENT device's stream: a tensor living on another CUDA device would be an illegal
    address inside the kernel, so refuse it here with a clear message."""
    cur = torch.cuda.current_device()
    for t in tensors:
        if t is not None and t.is_cuda and t.device.index != cur:
            raise RuntimeError(f"imageclassification_b200: tensor on {t.device} but the current CUDA device is cuda:{cur}; "
                               "call torch.cuda.set_device(...) (train.py:115 does) or wrap the call in torch.cuda.device(...)")


def require_cuda(*tensors, same_device: bool = True) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "imageclassification_b200 kernels run on CUDA (sm_100a) tensors only; got a tensor on "
                f"{t.device}. There is no CPU fallback — use oracle/ for CPU reference numbers.")
    if same_device:
        require_same_device(*tensors)
