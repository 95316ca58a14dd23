"""torch.autograd bindings of the libcnx kernels (one Function per reference op on the hot path).

Every Function launches hand-written sm_100a kernels through the C-ABI on the caller's current CUDA
stream; tensors (outputs, saved activations, scratch) are owned by PyTorch's caching allocator.
Reference ops replaced:
  block_forward        convnext.py:43-56 (Block.forward) == timm ConvNeXtBlock.forward
  layer_norm_cl        convnext.py:175-176 (LayerNorm, channels_last)
  soft_target_ce       timm SoftTargetCrossEntropy.forward (train.py:257)
  mixup_target         timm mixup_target (engine.py:44)
"""
from __future__ import annotations

import os
import weakref

import torch

from . import _lib as L

# Test-only switch: route the bf16 GEMMs through the CUDA-core kernel to cross-check the tcgen05 path.
GEMM_FLAGS = 0
# The fused no-grad MLP kernel (C <= 192); CNX_FUSED_MLP=0 in the environment selects the two-GEMM path for comparison.
FUSED_MLP = os.environ.get("CNX_FUSED_MLP", "1") != "0"
FUSED_MLP_X3 = os.environ.get("CNX_FUSED_MLP_X3", "1") != "0"     # fused split-operand MLP in the fp32 no-grad forward (C = 96)
# The reference's `gamma * x` (fp32 parameter x bf16 activation) and the residual add promote the stream to fp32 at the first
# Block after every bf16 downsample conv (convnext.py:52-55 under autocast).  CNX_BF16_STREAM=1 keeps the stream in bf16 through
# stages 1-3 instead (a third less activation traffic there, still inside the bf16 tolerance) — NOT the reference's dtypes,
# so it is opt-in and bench.py names it in `config` when set.
KEEP_BF16_STREAM = os.environ.get("CNX_BF16_STREAM", "0") == "1"
# fp32 (no autocast) no-grad forward — the reference's accuracy forward (engine.py:89-97) and evaluate() — on the tensor cores
# with split bf16 operands (include/cnx.h "x3"): fp32-accurate (~2^-16 per product), 3 MMAs per product instead of CUDA-core
# FMAs.  CNX_X3_FWD=0 selects the CUDA-core fp32 GEMMs for comparison.
X3_FWD = os.environ.get("CNX_X3_FWD", "1") != "0"
# Optional (CNX_RECOMPUTE_C=<max C>, default off): at C <= that, fc1 stores g only and the backward kernel recomputes the
# pre-activation for GELU' in a second TMEM accumulator (one [M,4C] tensor per Block instead of two; bit-identical results).
# Measured SLOWER on ConvNeXt-T/256 (profiles/r02k_*: fc1 K96 1.06 -> 0.89 ms but the data gradient 0.83 -> 1.19 ms; step 33.0 ->
# 33.6 ms): these K <= 192 GEMMs are bound by their epilogue pipeline (TMEM -> registers -> GELU math -> shared memory -> TMA
# store, ~2.3 us per 128 x 128 tile), not by the HBM writes the recomputation saves.  Kept for memory-limited runs: it saves
# M*4C*2 bytes of saved activations per Block (2.3 GB per step for ConvNeXt-T at batch 256).
RECOMPUTE_MAX_C = int(os.environ.get("CNX_RECOMPUTE_C", "0"))
# fp32 TRAINING (no autocast: the reference's --use_amp false default, BASELINE config 1) on the tensor cores: forward and the
# four backward GEMMs of the Block run as split-operand (x3) tcgen05 GEMMs, fp32-accurate to ~2^-16 per product, instead of the
# CUDA-core fp32 GEMMs.  CNX_X3_TRAIN=0 selects the CUDA-core kernels for comparison.
X3_TRAIN = os.environ.get("CNX_X3_TRAIN", "1") != "0"
X3_TRAIN_FUSED = os.environ.get("CNX_X3_TRAIN_FUSED", "1") != "0"   # fp32 training: GELU, split and GELU'(h) in the fc1 epilogue
# fp32 training weight gradients: one launch with a three-pass K loop instead of three (18 % faster, 3x the accumulation-chain
# error: 3-6e-5 against 1-2e-5 at the batch-256 shapes, include/cnx.h) — opt-in
WGRAD_X3_ONE_LOOP = os.environ.get("CNX_WGRAD_X3_ONE_LOOP", "0") == "1"


def _wgrad_x3_fn(lib):
    return lib.cnx_gemm_wgrad_x3_one_loop if WGRAD_X3_ONE_LOOP else lib.cnx_gemm_wgrad_x3
# Backward hand-off between consecutive Blocks (bf16 activations): the dwconv backward-data kernel of Block i writes, beside
# dx, the bf16 operand copy dz = bf16(dp * dx) that Block i-1's backward would make of it with cnx_grad_prep.  Forward notes
# which Block produced a Block's input (`_LAST_OUT`); backward leaves the copy in `_DZ_HANDOFF` together with the dx tensor
# it belongs to, and the consumer takes it only if the gradient it receives IS that tensor (same storage, shape, drop-path
# vector) — otherwise it falls back to cnx_grad_prep.  CNX_DZ_HANDOFF=0 disables it.
DZ_HANDOFF = os.environ.get("CNX_DZ_HANDOFF", "1") != "0"
_LAST_OUT = None        # (data_ptr, (N, H, W, C), dp) of the last Block output produced with grad tracking
_DZ_HANDOFF = None      # (dx NHWC tensor, dz [M, C] bf16, dp of the consumer)


def _act_dtype() -> torch.dtype:
    """bf16 activations under torch.autocast('cuda', dtype=torch.bfloat16), fp32 otherwise."""
    if torch.is_autocast_enabled("cuda"):
        d = torch.get_autocast_dtype("cuda")
        if d == torch.bfloat16:
            return torch.bfloat16
        raise NotImplementedError(
            f"imageclassification_b200 implements the bf16 autocast path (BASELINE north_star); autocast dtype {d} "
            "is not supported — use torch.autocast('cuda', dtype=torch.bfloat16)")
    return torch.float32


def _nhwc(x: torch.Tensor) -> torch.Tensor:
    """logical NCHW tensor -> contiguous [N,H,W,C] tensor (a view when x is channels-last strided)."""
    xl = x.permute(0, 2, 3, 1)
    return xl if xl.is_contiguous() else xl.contiguous()


def _num_partials(C: int = 1024) -> int:
    """CTAs (= partial rows) of the LayerNorm-backward kernels: 3 per SM for narrow rows, 2 per SM otherwise."""
    return L.load().cnx_sm_count() * (3 if C <= 256 else 2)


# Kernel-side layouts of the canonical fp32 parameters (bf16 copies, transposes, tap-major conv weights) are derived
# on the device and cached per parameter *version*: they are rebuilt only after the optimizer (or load_state_dict)
# has written the parameter, not on every forward / accuracy-forward / backward of the same step.
_DERIVED: dict = {}          # id(owner tensor) -> {tag: (key, derived tensor)}; entry dropped when the owner dies


def _derived(params, tag, build):
    """params: tuple of source tensors (first one owns the cache entry)."""
    owner = params[0]
    key = (tag,) + tuple((p.data_ptr(), p._version) for p in params if p is not None)
    slot = _DERIVED.get(id(owner))
    if slot is None:
        slot = {}
        _DERIVED[id(owner)] = slot
        weakref.finalize(owner, _DERIVED.pop, id(owner), None)
    hit = slot.get(tag)
    if hit is not None and hit[0] == key:
        return hit[1]
    val = build()
    slot[tag] = (key, val)
    return val


class _PrepRegistry:
    """Every derived GEMM weight layout (bf16 copy, transpose, gamma-scaled transpose) of the live models on one device.
    When any of them is found stale (the optimizer or load_state_dict wrote the parameters), ALL stale entries are rebuilt
    by ONE cnx_weight_prep_multi launch over a cached device table instead of one launch per weight."""

    def __init__(self, device):
        self.device = device
        self.entries = {}          # (id(w), mode, dtype) -> dict(w=weakref, scale=weakref|None, out, key)
        self.table = None          # (device table tensor, n, total_tiles, [entry keys in table order])
        self._keepalive = None

    def _drop(self, k):
        self.entries.pop(k, None)
        self.table = None

    @staticmethod
    def _key(w, scale):
        return (w.data_ptr(), w._version) + ((scale.data_ptr(), scale._version) if scale is not None else ())

    def get(self, w, mode, scale, out_dtype):
        k = (id(w), mode, out_dtype)
        e = self.entries.get(k)
        if e is None or e["w"]() is not w:
            R, Cc = w.shape[0], w.numel() // w.shape[0]          # [C,1,7,7] conv weights count as [C, 49]
            shape = (R, Cc) if mode == 0 else ((R, 3 * Cc) if mode == 3 else (Cc, R))
            out = torch.empty(shape, dtype=out_dtype, device=w.device)
            e = {"w": weakref.ref(w), "scale": weakref.ref(scale) if scale is not None else None, "out": out, "key": None,
                 "mode": mode, "dtype": out_dtype}
            self.entries[k] = e
            self.table = None
            weakref.finalize(w, self._drop, k)
            # a NEW layout is built on its own (one small launch); rebuilding every registered layout here made the first
            # forward of a model quadratic in its number of weights (128 multi-launches in the first step of ConvNeXt-T)
            L.check(L.load().cnx_weight_prep(L.ptr(w), R, Cc, L.ptr(scale), mode, L.ptr(out), L.dt(out_dtype), L.stream()),
                    "weight_prep")
            e["key"] = self._key(w, scale)
            return out
        if e["key"] != self._key(w, scale):
            self.refresh()
        return e["out"]

    def refresh(self):
        lib = L.load()
        live = []
        for k, e in list(self.entries.items()):
            w = e["w"]()
            sc = e["scale"]() if e["scale"] is not None else None
            if w is None or (e["scale"] is not None and sc is None):
                self._drop(k)
                continue
            live.append((k, e, w, sc))
        # entries of one source weight are adjacent and share their first tile: the kernel reads a source tile once and writes
        # every layout derived from it
        live.sort(key=lambda it: it[2].data_ptr())
        if self.table is None or self.table[3] != [(k, w.data_ptr(), sc.data_ptr() if sc is not None else 0) for k, e, w, sc in live]:
            rows, start, prev, ntile = [], 0, None, 0
            for k, e, w, sc in live:
                R, Cc = w.shape[0], w.numel() // w.shape[0]
                tx, ty = (Cc + 31) // 32, (R + 31) // 32
                if prev != (w.data_ptr(), R, Cc):
                    start += ntile
                    prev, ntile = (w.data_ptr(), R, Cc), tx * ty
                rows.append(L.WeightPrepEntry(w.data_ptr(), sc.data_ptr() if sc is not None else None, e["out"].data_ptr(), R, Cc,
                                              e["mode"], L.dt(e["dtype"]), start, tx))
            start += ntile
            arr = (L.WeightPrepEntry * len(rows))(*rows)
            host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
            dev = host.to(self.device)
            self.table = (dev, len(rows), start, [(k, w.data_ptr(), sc.data_ptr() if sc is not None else 0) for k, e, w, sc in live])
        dev, n, total, _ = self.table
        if n:
            L.check(lib.cnx_weight_prep_multi(L.ptr(dev), n, total, L.stream()), "weight_prep_multi")
        for k, e, w, sc in live:
            e["key"] = self._key(w, sc)


_PREP: dict = {}


def _weight_prep(w: torch.Tensor, mode: int, row_scale, out_dtype) -> torch.Tensor:
    if w.dtype != torch.float32 or not w.is_contiguous() or (row_scale is not None and row_scale.dtype != torch.float32):
        raise TypeError("libcnx keeps canonical parameters as contiguous fp32 tensors")
    reg = _PREP.get(w.device)
    if reg is None:
        reg = _PREP[w.device] = _PrepRegistry(w.device)
    return reg.get(w, mode, row_scale, out_dtype)


def _conv_weight_tap_major(conv_w: torch.Tensor) -> torch.Tensor:
    """[C,1,7,7] -> [49,C] fp32 (what the dwconv kernels fetch by TMA): a transpose of the [C,49] matrix, refreshed together
    with every other derived weight layout by the one cnx_weight_prep_multi launch per optimizer step."""
    if conv_w.dtype == torch.float32 and conv_w.is_contiguous() and conv_w.shape[1:] == (1, 7, 7):
        return _weight_prep(conv_w, 1, None, torch.float32)

    def build():
        lib = L.load()
        C = conv_w.shape[0]
        wt = torch.empty((49, C), dtype=torch.float32, device=conv_w.device)
        L.check(lib.cnx_dwconv7_weight_prep(L.ptr(conv_w), C, L.ptr(wt), L.stream()), "dwconv7_weight_prep")
        return wt
    return _derived((conv_w,), ("tapmajor",), build)


# Gradient sinks: a data-parallel reducer (ddp.DistributedDataParallel) registers, per parameter, an object that names the
# buffer this parameter's gradient should be written into (a slot of its flat arena).  The backward functions below then point
# their weight-gradient kernels' `out` / `accumulate` arguments at that slot, return None for the parameter (autograd neither
# allocates nor copies anything) and report the parameter ready.  Kept out of the parameters' __dict__ on purpose: whole-model
# pickling (utils.py:542) serialises parameter attributes.
_GRAD_SINKS: dict = {}          # id(param) -> (weakref to param, sink); tensors cannot key a WeakKeyDictionary (== is elementwise)


def register_grad_sink(param: torch.Tensor, sink) -> None:
    key = id(param)
    _GRAD_SINKS[key] = (weakref.ref(param), sink)
    weakref.finalize(param, _GRAD_SINKS.pop, key, None)


def _sink_of(param):
    if param is None or not param.requires_grad:
        return None
    e = _GRAD_SINKS.get(id(param))
    return e[1] if (e is not None and e[0]() is param) else None


class _Dest:
    """Destination of one kernel's group of parameter gradients (e.g. fc1 weight + bias): either every parameter's arena slot
    (`sunk`), or fresh tensors that are returned to autograd.  One accumulate flag per group, as the kernels have."""
    __slots__ = ("bufs", "acc", "sink", "params")

    def __init__(self, params, shapes):
        self.params = params
        self.sink, self.acc = None, 0
        sinks = [_sink_of(p) for p in params]
        if sinks[0] is not None and all(s is sinks[0] for s in sinks):
            claims = [sinks[0].claim(p) for p in params]
            if all(c is not None for c in claims) and len({c[1] for c in claims}) == 1:
                self.bufs = [c[0] for c in claims]
                self.acc = claims[0][1]
                self.sink = sinks[0]
                return
        dev = next(p for p in params if p is not None).device
        self.bufs = [torch.empty(sh, dtype=torch.float32, device=dev) if p is not None else None for p, sh in zip(params, shapes)]

    def results(self):
        """after the kernels were enqueued: what backward returns for these parameters"""
        if self.sink is None:
            return self.bufs
        for p in self.params:
            self.sink.ready(p)
        return [None] * len(self.bufs)


def _wgrad(X, Y, M, N1, N2, want_colsum: bool, out=None, cs=None, accumulate: int = 0):
    lib = L.load()
    if X.dtype == torch.float32 and Y.dtype == torch.float32 and X3_TRAIN and GEMM_FLAGS == 0 and N1 % 8 == 0 and N2 % 8 == 0:
        # fp32 operands: three bf16 tensor-core wgrad GEMMs over the split operands (fp32-accurate), not the CUDA-core kernel
        bf = torch.bfloat16
        x2 = torch.empty((M, 2 * N1), dtype=bf, device=X.device)
        y2 = torch.empty((M, 2 * N2), dtype=bf, device=X.device)
        L.check(lib.cnx_split3(L.ptr(X), M, N1, L.ptr(x2), 2, L.stream()), "split3")
        L.check(lib.cnx_split3(L.ptr(Y), M, N2, L.ptr(y2), 2, L.stream()), "split3")
        ws_bytes = lib.cnx_gemm_wgrad_workspace_bytes(M, N1, N2, L.CNX_BF16, 0)
        ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=X.device)
        if out is None:
            out = torch.empty((N1, N2), dtype=torch.float32, device=X.device)
        if cs is None and want_colsum:
            cs = torch.empty((N1,), dtype=torch.float32, device=X.device)
        L.check(_wgrad_x3_fn(lib)(L.ptr(x2), L.ptr(y2), M, N1, N2, int(accumulate), L.ptr(out), L.ptr(cs), L.ptr(ws), ws_bytes,
                                      L.stream()), "gemm_wgrad_x3")
        return out, cs
    d = L.dt(X)
    ws_bytes = lib.cnx_gemm_wgrad_workspace_bytes(M, N1, N2, d, GEMM_FLAGS)
    ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=X.device)
    if out is None:
        out = torch.empty((N1, N2), dtype=torch.float32, device=X.device)
    if cs is None and want_colsum:
        cs = torch.empty((N1,), dtype=torch.float32, device=X.device)
    L.check(lib.cnx_gemm_wgrad(L.ptr(X), L.ptr(Y), M, N1, N2, int(accumulate), L.ptr(out), L.ptr(cs), L.ptr(ws), ws_bytes, d,
                               GEMM_FLAGS, L.stream()), "gemm_wgrad")
    return out, cs


def _split_of(params, tag, build32, rows, cols):
    """[hi | hi | mid] bf16 split (weight-prep mode 3) of a derived fp32 weight layout, rebuilt when the parameters change"""
    def build():
        w32 = build32()
        out = torch.empty((rows, 3 * cols), dtype=torch.bfloat16, device=w32.device)
        L.check(L.load().cnx_weight_prep(L.ptr(w32), rows, cols, None, 3, L.ptr(out), L.dt(torch.bfloat16), L.stream()), "weight_prep(x3)")
        return out
    return _derived(params, tag, build)


def _mlp_backward_x3(lib, params, doutl, a2, gp, g2, w1, w2, b2, gamma, dp, M, C, C4, rows_per_sample, dev, st):
    """Backward of fc1 -> GELU -> fc2 -> layer-scale in fp32 accuracy on the tensor cores (split operands):
    -> dxn fp32 [M,C], dW1, db1, dW2, db2, dgamma (None where a gradient sink took the result)."""
    p_w1, p_b1, p_w2, p_b2, p_gamma = params[4:9]
    bf, f32 = torch.bfloat16, torch.float32
    # 1. dz = dp * dout (fp32), as the split operand [hi | mid]
    dz32 = doutl.reshape(M, C)
    if dp is not None:
        t = torch.empty((M, C), dtype=f32, device=dev)
        L.check(lib.cnx_grad_prep(L.ptr(dz32), L.dt(f32), L.ptr(dp), rows_per_sample, M, C, L.ptr(t), L.dt(f32), st), "grad_prep")
        dz32 = t
    dz2 = torch.empty((M, 2 * C), dtype=bf, device=dev)
    L.check(lib.cnx_split3(L.ptr(dz32), M, C, L.ptr(dz2), 2, st), "split3")
    # 2. dh = (dz . (gamma*W2)) * GELU'(h), leaving as a split operand
    w2gt3 = _split_of((w2, gamma), ("w2gt_x3",), lambda: _weight_prep(w2, 2, gamma, f32), C4, C)            # [4C, 3C]
    dh2 = torch.empty((M, 2 * C4), dtype=bf, device=dev)
    if X3_TRAIN_FUSED and C4 % 32 == 0:
        # the multiply by GELU'(h) and the hi / mid split happen in the GEMM's epilogue: no fp32 [M, 4C] product round trip
        L.check(lib.cnx_gemm_dgrad_gelu_bwd_x3(L.ptr(dz2), L.ptr(w2gt3), L.ptr(gp), L.ptr(dh2), M, C4, 3 * C, 2, st),
                "gemm_dgrad_gelu_bwd_x3")
    else:
        t1 = torch.empty((M, C4), dtype=f32, device=dev)
        L.check(lib.cnx_gemm_plain(L.ptr(dz2), L.ptr(w2gt3), None, L.ptr(t1), L.dt(f32), M, C4, 3 * C, L.dt(bf), L.CNX_GEMM_A_SPLIT2, st),
                "gemm_plain(x3 dgrad fc2)")
        L.check(lib.cnx_mul_split(L.ptr(t1), L.ptr(gp), M, C4, L.ptr(dh2), st), "mul_split")
        del t1
    # 3. fc2 wgrad on the unscaled gradient + layer-scale identities
    ws_bytes = max(lib.cnx_gemm_wgrad_workspace_bytes(M, C, C4, L.CNX_BF16, 0), lib.cnx_gemm_wgrad_workspace_bytes(M, C4, C, L.CNX_BF16, 0))
    ws = torch.empty(ws_bytes // 4, dtype=f32, device=dev)
    G2 = torch.empty((C, C4), dtype=f32, device=dev)
    s = torch.empty((C,), dtype=f32, device=dev)
    L.check(_wgrad_x3_fn(lib)(L.ptr(dz2), L.ptr(g2), M, C, C4, 0, L.ptr(G2), L.ptr(s), L.ptr(ws), ws_bytes, st), "gemm_wgrad_x3")
    d_fc2 = _Dest((p_w2, p_b2, p_gamma), (w2.shape, b2.shape, (C,)))
    L.check(lib.cnx_layerscale_finalize(L.ptr(G2), L.ptr(s), L.ptr(w2), L.ptr(b2), L.ptr(gamma), C, C4, d_fc2.acc,
                                        L.ptr(d_fc2.bufs[0]), L.ptr(d_fc2.bufs[1]), L.ptr(d_fc2.bufs[2]), st), "layerscale_finalize")
    dW2, db2, dgamma = d_fc2.results()
    # 4. dxn = dh . W1
    w1t3 = _split_of((w1,), ("w1t_x3",), lambda: _weight_prep(w1, 1, None, f32), C, C4)                      # [C, 3*4C]
    dxn = torch.empty((M, C), dtype=f32, device=dev)
    L.check(lib.cnx_gemm_plain(L.ptr(dh2), L.ptr(w1t3), None, L.ptr(dxn), L.dt(f32), M, C, 3 * C4, L.dt(bf), L.CNX_GEMM_A_SPLIT2, st),
            "gemm_plain(x3 dgrad fc1)")
    # 5. fc1 wgrad + bias grad
    d_fc1 = _Dest((p_w1, p_b1), (w1.shape, (C4,)))
    L.check(_wgrad_x3_fn(lib)(L.ptr(dh2), L.ptr(a2), M, C4, C, d_fc1.acc, L.ptr(d_fc1.bufs[0]), L.ptr(d_fc1.bufs[1]), L.ptr(ws),
                                  ws_bytes, st), "gemm_wgrad_x3")
    dW1, db1 = d_fc1.results()
    return dxn, dW1, db1, dW2, db2, dgamma


class _BlockFn(torch.autograd.Function):
    """dwconv7 -> LN -> fc1 -> GELU -> fc2 -> gamma -> drop_path -> + shortcut, forward and backward."""

    @staticmethod
    def forward(ctx, x, conv_w, conv_b, ln_w, ln_b, w1, b1, w2, b2, gamma, dp, eps, act_dtype, track):
        global _LAST_OUT, _DZ_HANDOFF
        last, _LAST_OUT = _LAST_OUT, None
        lib = L.load()
        L.require_cuda(x, conv_w, w1, w2)
        if x.dtype not in (torch.float32, torch.bfloat16):
            raise TypeError(f"ConvNeXtBlock: unsupported residual-stream dtype {x.dtype}")
        N, C, H, W = x.shape
        xl = _nhwc(x.detach())
        M = N * H * W
        dev = x.device
        sd, ad = L.dt(xl), L.dt(act_dtype)
        st = L.stream()
        C4 = w1.shape[0]
        # `track` = grad mode of the CALLER (inside Function.forward grad mode is always off; needs_input_grad alone is True
        # for parameters even under torch.no_grad())
        need_grad = track and (ctx.needs_input_grad[0] or any(ctx.needs_input_grad[1:10]))
        y = torch.empty((M, C), dtype=act_dtype, device=dev)
        mean = torch.empty((M,), dtype=torch.float32, device=dev)
        rstd = torch.empty((M,), dtype=torch.float32, device=dev)
        wt = _conv_weight_tap_major(conv_w)
        if not need_grad and X3_FWD and act_dtype == torch.float32 and C % 32 == 0 and C4 % 64 == 0:
            # fp32 no-grad pass on the tensor cores: split operands [hi | mid | hi] x [hi | hi | mid], K' = 3K; the LayerNorm half
            # of the dwconv kernel writes the split operand itself
            bf = torch.bfloat16
            a3 = torch.empty((M, 2 * C), dtype=bf, device=dev)             # [hi | mid] of xn: fc1's K loop wraps for the third segment
            L.check(lib.cnx_dwconv7_ln_fwd_x3(L.ptr(xl), L.ptr(wt), L.ptr(conv_b), L.ptr(ln_w), L.ptr(ln_b), eps, N, H, W, C,
                                              L.ptr(y), L.ptr(a3), L.ptr(mean), L.ptr(rstd), 2, st), "dwconv7_ln_fwd_x3")
            out = torch.empty((N, H, W, C), dtype=xl.dtype, device=dev)
            if FUSED_MLP_X3 and C == 96 and C4 == 4 * C and M >= 128:
                # fc1 -> GELU -> split -> fc2 -> gamma / drop-path / residual in one kernel: the hidden activation (85 % of the
                # bytes the unfused pair moves at C = 96) never reaches HBM
                L.check(lib.cnx_mlp_fused_fwd_x3(L.ptr(a3), L.ptr(_weight_prep(w1, 3, None, bf)), L.ptr(b1),
                                                 L.ptr(_weight_prep(w2, 3, None, bf)), L.ptr(b2), L.ptr(gamma), L.ptr(dp), H * W,
                                                 L.ptr(xl), L.ptr(out), M, C, st), "mlp_fused_fwd_x3")
                return out.permute(0, 3, 1, 2)
            g2 = torch.empty((M, 2 * C4), dtype=bf, device=dev)            # [hi | mid] of g
            L.check(lib.cnx_gemm_bias_gelu_fwd_x3(L.ptr(a3), L.ptr(_weight_prep(w1, 3, None, bf)), L.ptr(b1), M, C4, 3 * C,
                                                  L.ptr(g2), 2, st), "gemm_bias_gelu_fwd_x3")
            del a3
            # (running fc1 -> fc2 over L2-sized row blocks so that g never reaches HBM was measured SLOWER: 36.9-41.1 ms/step
            # against 35.0 for blocks of 96-32 MB — launch-bound, and the L2 does not keep a written block resident)
            L.check(lib.cnx_gemm_bias_scale_residual_fwd(L.ptr(g2), L.ptr(_weight_prep(w2, 3, None, bf)), L.ptr(b2), L.ptr(gamma),
                                                         L.ptr(dp), H * W, L.ptr(xl), L.ptr(out), sd, M, C, 3 * C4, L.dt(bf),
                                                         L.CNX_GEMM_A_SPLIT2, st), "gemm_bias_scale_residual_fwd(x3)")
            return out.permute(0, 3, 1, 2)
        if need_grad and X3_TRAIN and act_dtype == torch.float32 and C % 32 == 0 and C4 % 64 == 0 and GEMM_FLAGS == 0:
            # fp32 training on the tensor cores: every GEMM operand crosses as [hi | mid] bf16 pieces and the K loops cover
            # hi.hi + mid.hi + hi.mid (include/cnx.h "x3").  Saved for backward: y, the split xn, GELU'(h) (fp32, in h's buffer),
            # the split g
            bf = torch.bfloat16
            a2 = torch.empty((M, 2 * C), dtype=bf, device=dev)
            L.check(lib.cnx_dwconv7_ln_fwd_x3(L.ptr(xl), L.ptr(wt), L.ptr(conv_b), L.ptr(ln_w), L.ptr(ln_b), eps, N, H, W, C,
                                              L.ptr(y), L.ptr(a2), L.ptr(mean), L.ptr(rstd), 2, st), "dwconv7_ln_fwd_x3")
            h = torch.empty((M, C4), dtype=torch.float32, device=dev)
            g2 = torch.empty((M, 2 * C4), dtype=bf, device=dev)
            if X3_TRAIN_FUSED and C4 % 32 == 0:
                # fc1 + bias + erf GELU + hi / mid split + GELU'(h) from ONE epilogue: no fp32 pre-activation round trip
                L.check(lib.cnx_gemm_bias_gelu_fwd_x3_train(L.ptr(a2), L.ptr(_weight_prep(w1, 3, None, bf)), L.ptr(b1), M, C4, 3 * C,
                                                            L.ptr(g2), L.ptr(h), 2, st), "gemm_bias_gelu_fwd_x3_train")
            else:
                L.check(lib.cnx_gemm_plain(L.ptr(a2), L.ptr(_weight_prep(w1, 3, None, bf)), L.ptr(b1), L.ptr(h), L.dt(torch.float32), M,
                                           C4, 3 * C, L.dt(bf), L.CNX_GEMM_A_SPLIT2, st), "gemm_plain(x3 fc1)")
                L.check(lib.cnx_gelu_split(L.ptr(h), M, C4, L.ptr(g2), st), "gelu_split")     # h now holds GELU'(h)
            out = torch.empty((N, H, W, C), dtype=xl.dtype, device=dev)
            L.check(lib.cnx_gemm_bias_scale_residual_fwd(L.ptr(g2), L.ptr(_weight_prep(w2, 3, None, bf)), L.ptr(b2), L.ptr(gamma),
                                                         L.ptr(dp), H * W, L.ptr(xl), L.ptr(out), sd, M, C, 3 * C4, L.dt(bf),
                                                         L.CNX_GEMM_A_SPLIT2, st), "gemm_bias_scale_residual_fwd(x3)")
            ctx.save_for_backward(xl, y, a2, mean, rstd, h, g2, conv_w, ln_w, w1, w2, b2, gamma, dp)
            ctx.shape = (N, C, H, W)
            ctx.act_dtype = act_dtype
            ctx.x3 = True
            ctx.up_dp = None
            ctx.params = (conv_w, conv_b, ln_w, ln_b, w1, b1, w2, b2, gamma)
            return out.permute(0, 3, 1, 2)
        xn = torch.empty((M, C), dtype=act_dtype, device=dev)
        L.check(lib.cnx_dwconv7_ln_fwd(L.ptr(xl), sd, L.ptr(wt), L.ptr(conv_b), L.ptr(ln_w), L.ptr(ln_b), eps,
                                       N, H, W, C, L.ptr(y), L.ptr(xn), ad, L.ptr(mean), L.ptr(rstd), st), "dwconv7_ln_fwd")
        if act_dtype == torch.float32:
            w1a, w2a = w1, w2
        else:
            w1a = _weight_prep(w1, 0, None, act_dtype)
            w2a = _weight_prep(w2, 0, None, act_dtype)
        if (not need_grad and FUSED_MLP and act_dtype == torch.bfloat16 and xl.dtype == torch.float32 and C in (96, 128, 192)
                and C4 == 4 * C and M >= 128):
            # no-grad pass (engine.py:89-97 accuracy forward, evaluate()): fc1 -> GELU -> fc2 -> gamma / drop-path / residual
            # in one kernel, the hidden activation never reaches HBM
            out = torch.empty((N, H, W, C), dtype=xl.dtype, device=dev)
            L.check(lib.cnx_mlp_fused_fwd(L.ptr(xn), L.ptr(w1a), L.ptr(b1), L.ptr(w2a), L.ptr(b2), L.ptr(gamma), L.ptr(dp), H * W,
                                          L.ptr(xl), L.ptr(out), M, C, st), "mlp_fused_fwd")
            return out.permute(0, 3, 1, 2)
        up = None
        if need_grad and DZ_HANDOFF and act_dtype == torch.bfloat16:
            _DZ_HANDOFF = None
            if last is not None and last[0] == xl.data_ptr() and last[1] == (N, H, W, C) and ctx.needs_input_grad[0]:
                up = last                            # this Block's input is the previous Block's output
        recompute = (need_grad and act_dtype == torch.bfloat16 and C <= RECOMPUTE_MAX_C and C4 % 128 == 0 and M >= 256
                     and GEMM_FLAGS == 0)
        h = torch.empty((M, C4), dtype=act_dtype, device=dev) if (need_grad and not recompute) else None     # holds GELU'(h)
        g = torch.empty((M, C4), dtype=act_dtype, device=dev)
        L.check(lib.cnx_gemm_bias_gelu_fwd(L.ptr(xn), L.ptr(w1a), L.ptr(b1), M, C4, C, L.ptr(h), L.ptr(g), ad,
                                           GEMM_FLAGS, st), "gemm_bias_gelu_fwd")
        out = torch.empty((N, H, W, C), dtype=xl.dtype, device=dev)
        L.check(lib.cnx_gemm_bias_scale_residual_fwd(L.ptr(g), L.ptr(w2a), L.ptr(b2), L.ptr(gamma), L.ptr(dp), H * W,
                                                     L.ptr(xl), L.ptr(out), sd, M, C, C4, ad, GEMM_FLAGS, st),
                "gemm_bias_scale_residual_fwd")
        if need_grad:
            ctx.save_for_backward(xl, y, xn, mean, rstd, h, g, conv_w, ln_w, w1, w2, b2, gamma, dp)
            ctx.shape = (N, C, H, W)
            ctx.act_dtype = act_dtype
            ctx.x3 = False
            ctx.params = (conv_w, conv_b, ln_w, ln_b, w1, b1, w2, b2, gamma)     # the leaves themselves (gradient sinks)
            ctx.up_dp = (up[2],) if up is not None else None                      # drop-path vector of the producing Block
            if DZ_HANDOFF and act_dtype == torch.bfloat16:
                _LAST_OUT = (out.data_ptr(), (N, H, W, C), dp)
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dout):
        global _DZ_HANDOFF
        ho, _DZ_HANDOFF = _DZ_HANDOFF, None
        lib = L.load()
        xl, y, xn, mean, rstd, h, g, conv_w, ln_w, w1, w2, b2, gamma, dp = ctx.saved_tensors
        N, C, H, W = ctx.shape
        act_dtype = ctx.act_dtype
        M, C4 = N * H * W, w1.shape[0]
        dev = xl.device
        sd, ad = L.dt(xl), L.dt(act_dtype)
        st = L.stream()
        doutl = _nhwc(dout)
        if doutl.dtype != xl.dtype:
            doutl = doutl.to(xl.dtype)
        p_conv_w, p_conv_b, p_ln_w, p_ln_b, p_w1, p_b1, p_w2, p_b2, p_gamma = ctx.params
        if ctx.x3:
            ho = None
            dxn, dW1, db1, dW2, db2, dgamma = _mlp_backward_x3(lib, ctx.params, doutl, xn, h, g, w1, w2, b2, gamma, dp, M, C, C4,
                                                              H * W, dev, st)
        else:
            # 1. dz = act(dp * dout): the operand copy of the incoming gradient (drop-path folded in once)
            if dp is None and act_dtype == xl.dtype:
                dz = doutl.reshape(M, C)
            elif (ho is not None and ho[0].data_ptr() == doutl.data_ptr() and ho[0].shape == doutl.shape and ho[0].dtype == doutl.dtype
                  and ((ho[2] is None) if dp is None else (ho[2] is not None and ho[2].data_ptr() == dp.data_ptr()))
                  and act_dtype == torch.bfloat16):
                dz = ho[1]                           # written by the downstream Block's dwconv backward-data kernel
            else:
                dz = torch.empty((M, C), dtype=act_dtype, device=dev)
                L.check(lib.cnx_grad_prep(L.ptr(doutl), sd, L.ptr(dp), H * W, M, C, L.ptr(dz), ad, st), "grad_prep")
            # 2. dh = (dz . (gamma*W2)) * GELU'(h)
            w2gt = _weight_prep(w2, 2, gamma, act_dtype)            # [4C, C]
            dh = torch.empty((M, C4), dtype=act_dtype, device=dev)
            if h is None:
                # GELU'(h) was not saved: the kernel recomputes h = xn . W1^T + b1 beside the data gradient
                L.check(lib.cnx_gemm_dgrad_gelu_recompute_bwd(L.ptr(dz), L.ptr(w2gt), L.ptr(xn), L.ptr(_weight_prep(w1, 0, None, act_dtype)),
                                                              L.ptr(ctx.params[5]), L.ptr(dh), M, C4, C, ad, st),
                        "gemm_dgrad_gelu_recompute_bwd")
            else:
                L.check(lib.cnx_gemm_dgrad_gelu_bwd(L.ptr(dz), L.ptr(w2gt), L.ptr(h), L.ptr(dh), M, C4, C, ad, GEMM_FLAGS, st),
                        "gemm_dgrad_gelu_bwd")
            # 3. fc2 wgrad on the UNSCALED gradient; layer-scale identities give dW2, db2, dgamma without saving z
            G2, s = _wgrad(dz, g, M, C, C4, True)
            d_fc2 = _Dest((p_w2, p_b2, p_gamma), (w2.shape, b2.shape, (C,)))
            dW2, db2, dgamma = d_fc2.bufs
            L.check(lib.cnx_layerscale_finalize(L.ptr(G2), L.ptr(s), L.ptr(w2), L.ptr(b2), L.ptr(gamma), C, C4, d_fc2.acc,
                                                L.ptr(dW2), L.ptr(db2), L.ptr(dgamma), st), "layerscale_finalize")
            dW2, db2, dgamma = d_fc2.results()
            # 4. dxn = dh . W1
            w1t = _weight_prep(w1, 1, None, act_dtype)              # [C, 4C]
            dxn = torch.empty((M, C), dtype=act_dtype, device=dev)
            L.check(lib.cnx_gemm_plain(L.ptr(dh), L.ptr(w1t), None, L.ptr(dxn), ad, M, C, C4, ad, GEMM_FLAGS, st), "gemm_plain")
            # 5. fc1 wgrad + bias grad
            d_fc1 = _Dest((p_w1, p_b1), (w1.shape, (C4,)))
            _wgrad(dh, xn, M, C4, C, True, out=d_fc1.bufs[0], cs=d_fc1.bufs[1], accumulate=d_fc1.acc)
            dW1, db1 = d_fc1.results()
        # 6. LayerNorm backward
        P = _num_partials(C)
        dy = torch.empty((M, C), dtype=act_dtype, device=dev)
        part = torch.empty((P, 2 * C), dtype=torch.float32, device=dev)
        L.check(lib.cnx_ln_bwd(L.ptr(dxn), ad, L.ptr(y), ad, L.ptr(mean), L.ptr(rstd), L.ptr(ln_w), M, C, L.ptr(dy), ad,
                               L.ptr(part), P, st), "ln_bwd")
        d_ln = _Dest((p_ln_w, p_ln_b), ((C,), (C,)))
        L.check(lib.cnx_reduce_partials_split(L.ptr(part), P, C, C, d_ln.acc, L.ptr(d_ln.bufs[0]), L.ptr(d_ln.bufs[1]), st),
                "reduce_partials_split")
        dln_w, dln_b = d_ln.results()
        # 7. dwconv wgrad (+bias)
        # one persistent CTA per SM: (C/32 channel chunks) x Pw partial rows ~= SM count, every CTA sweeps many tiles
        Pw = max(1, min(L.load().cnx_sm_count() // max(C // 32, 1), (N * ((H + 7) // 8) * ((W + 31) // 32))))
        wpart = torch.empty((Pw, 50, C), dtype=torch.float32, device=dev)
        L.check(lib.cnx_dwconv7_wgrad(L.ptr(dy), ad, L.ptr(xl), sd, N, H, W, C, L.ptr(wpart), Pw, st), "dwconv7_wgrad")
        d_cv = _Dest((p_conv_w, p_conv_b), (conv_w.shape, (C,)))
        L.check(lib.cnx_dwconv7_wgrad_finalize(L.ptr(wpart), Pw, C, d_cv.acc, L.ptr(d_cv.bufs[0]), L.ptr(d_cv.bufs[1]), st),
                "dwconv7_wgrad_finalize")
        dconv_w, dconv_b = d_cv.results()
        # 8. dx = dout + dwconv_dgrad(dy)
        dx = None
        if ctx.needs_input_grad[0]:
            dxl = torch.empty((N, H, W, C), dtype=xl.dtype, device=dev)
            if ctx.up_dp is not None and not ctx.x3 and act_dtype == torch.bfloat16:
                dz_up = torch.empty((M, C), dtype=act_dtype, device=dev)
                L.check(lib.cnx_dwconv7_dgrad_dz(L.ptr(dy), L.ptr(_conv_weight_tap_major(conv_w)), L.ptr(doutl), L.ptr(dxl), sd, N, H, W, C,
                                                 L.ptr(dz_up), L.ptr(ctx.up_dp[0]), st), "dwconv7_dgrad_dz")
                _DZ_HANDOFF = (dxl, dz_up, ctx.up_dp[0])
            else:
                L.check(lib.cnx_dwconv7_dgrad(L.ptr(dy), ad, L.ptr(_conv_weight_tap_major(conv_w)), L.ptr(doutl), L.ptr(dxl), sd, N, H, W, C, st),
                        "dwconv7_dgrad")
            dx = dxl.permute(0, 3, 1, 2)
        return (dx, dconv_w, dconv_b, dln_w, dln_b, dW1, db1, dW2, db2, dgamma, None, None, None, None)


def block_forward(x, conv_w, conv_b, ln_w, ln_b, w1, b1, w2, b2, gamma, dp, eps: float):
    """ConvNeXt Block forward (autograd-enabled).  x: logical [N,C,H,W]; dp: per-sample drop-path scale [N] or None."""
    act_dtype = _act_dtype()
    if act_dtype == torch.float32 and x.dtype != torch.float32:
        raise TypeError("fp32 mode (no autocast) needs an fp32 residual stream")
    if x.dtype == torch.bfloat16 and gamma is not None and gamma.dtype == torch.float32 and not KEEP_BF16_STREAM:
        x = x.float()          # same values; the Block then returns fp32 exactly as ATen's type promotion does in the reference
    return _BlockFn.apply(x, conv_w.contiguous(), conv_b, ln_w, ln_b, w1, b1, w2, b2, gamma, dp, float(eps), act_dtype,
                          torch.is_grad_enabled())


class _LayerNormCLFn(torch.autograd.Function):
    """LayerNorm over the last dim of a contiguous [..., C] tensor; fp32 output under autocast (as ATen's policy)."""

    @staticmethod
    def forward(ctx, x, w, b, eps, out_dtype):
        lib = L.load()
        L.require_cuda(x, w, b)
        C = x.shape[-1]
        x2 = x.detach().reshape(-1, C)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        M = x2.shape[0]
        out = torch.empty((M, C), dtype=out_dtype, device=x.device)
        mean = torch.empty((M,), dtype=torch.float32, device=x.device)
        rstd = torch.empty((M,), dtype=torch.float32, device=x.device)
        L.check(lib.cnx_ln_fwd(L.ptr(x2), L.dt(x2), L.ptr(w), L.ptr(b), eps, M, C, L.ptr(out), L.dt(out_dtype),
                               L.ptr(mean), L.ptr(rstd), L.stream()), "ln_fwd")
        ctx.save_for_backward(x2, mean, rstd, w)
        ctx.xshape = x.shape
        return out.reshape(x.shape)

    @staticmethod
    def backward(ctx, dout):
        lib = L.load()
        x2, mean, rstd, w = ctx.saved_tensors
        M, C = x2.shape
        d2 = dout.reshape(M, C)
        if not d2.is_contiguous():
            d2 = d2.contiguous()
        P = max(1, min(_num_partials(C), (M + 7) // 8))
        dx = torch.empty((M, C), dtype=x2.dtype, device=x2.device)
        part = torch.empty((P, 2 * C), dtype=torch.float32, device=x2.device)
        L.check(lib.cnx_ln_bwd(L.ptr(d2), L.dt(d2), L.ptr(x2), L.dt(x2), L.ptr(mean), L.ptr(rstd), L.ptr(w), M, C,
                               L.ptr(dx), L.dt(dx), L.ptr(part), P, L.stream()), "ln_bwd")
        dwb = torch.empty((2 * C,), dtype=torch.float32, device=x2.device)
        L.check(lib.cnx_reduce_partials(L.ptr(part), P, 2 * C, 1.0, 0, L.ptr(dwb), L.stream()), "reduce_partials")
        return dx.reshape(ctx.xshape), dwb[:C], dwb[C:], None, None


def layer_norm_cl(x, w, b, eps: float):
    """F.layer_norm(x, (C,), w, b, eps) for channels-last x; output fp32 under autocast (ATen autocast policy)."""
    out_dtype = torch.float32 if (torch.is_autocast_enabled("cuda") or x.dtype == torch.float32) else x.dtype
    return _LayerNormCLFn.apply(x, w, b, float(eps), out_dtype)


def _gemm_plain(A, B, bias, out_dtype, round_bf16=False):
    """out[M,N] = A[M,K] . B[N,K]^T (+ bias[N]); A, B in the activation dtype, fp32 accumulate.  round_bf16: an fp32 output that
    holds bf16-rounded values (include/cnx.h CNX_GEMM_OUT_ROUND_BF16)."""
    lib = L.load()
    M, K = A.shape
    Nn = B.shape[0]
    if (A.dtype == torch.float32 and B.dtype == torch.float32 and out_dtype == torch.float32 and X3_TRAIN and GEMM_FLAGS == 0
            and K % 8 == 0 and Nn % 8 == 0):
        return _gemm_plain_f32_x3(A, B, bias)          # fp32 operands: fp32-accurate split-operand GEMM on the tensor cores
    out = torch.empty((M, Nn), dtype=out_dtype, device=A.device)
    L.check(lib.cnx_gemm_plain(L.ptr(A), L.ptr(B), L.ptr(bias), L.ptr(out), L.dt(out_dtype), M, Nn, K, L.dt(A),
                               GEMM_FLAGS | (L.CNX_GEMM_OUT_ROUND_BF16 if round_bf16 else 0), L.stream()), "gemm_plain")
    return out


def _gemm_plain_f32_x3(A, B32, bias):
    """fp32 out[M,N] = A[M,K] . B32[N,K]^T + bias with split bf16 operands: B's [hi | hi | mid] layout is cached per version of
    B32, A is split on the fly ([hi | mid] + wrapping K loop where K is a multiple of 32, three segments otherwise)."""
    lib = L.load()
    M, K = A.shape
    Nn = B32.shape[0]
    bf = torch.bfloat16

    def build():
        out = torch.empty((Nn, 3 * K), dtype=bf, device=B32.device)
        L.check(lib.cnx_weight_prep(L.ptr(B32), Nn, K, None, 3, L.ptr(out), L.dt(bf), L.stream()), "weight_prep(x3)")
        return out
    B3 = _derived((B32,), ("b_x3",), build)
    seg2 = K % 32 == 0
    a3 = torch.empty((M, (2 if seg2 else 3) * K), dtype=bf, device=A.device)
    L.check(lib.cnx_split3(L.ptr(A), M, K, L.ptr(a3), 2 if seg2 else 3, L.stream()), "split3")
    out = torch.empty((M, Nn), dtype=torch.float32, device=A.device)
    L.check(lib.cnx_gemm_plain(L.ptr(a3), L.ptr(B3), L.ptr(bias), L.ptr(out), L.dt(torch.float32), M, Nn, 3 * K, L.dt(bf),
                               L.CNX_GEMM_A_SPLIT2 if seg2 else 0, L.stream()), "gemm_plain(x3)")
    return out


def _gemm_plain_x3(A, conv_w, channels_last_taps: bool, bias):
    """fp32 out[M,N] = A[M,K] . W[N,K]^T + bias on the tensor cores with split bf16 operands (fp32-accurate, no-grad forward)."""
    lib = L.load()
    M, K = A.shape
    bf = torch.bfloat16

    def build():
        w2 = _patch_weight(conv_w, torch.float32, channels_last_taps)
        out = torch.empty((w2.shape[0], 3 * w2.shape[1]), dtype=bf, device=w2.device)
        L.check(lib.cnx_weight_prep(L.ptr(w2), w2.shape[0], w2.shape[1], None, 3, L.ptr(out), L.dt(bf), L.stream()), "weight_prep(x3)")
        return out
    B3 = _derived((conv_w,), ("patchw_x3", channels_last_taps), build)
    seg2 = K % 32 == 0                                         # two segments + a wrapping K loop where the segment is whole 16-column MMA slices
    a3 = torch.empty((M, (2 if seg2 else 3) * K), dtype=bf, device=A.device)
    L.check(lib.cnx_split3(L.ptr(A), M, K, L.ptr(a3), 2 if seg2 else 3, L.stream()), "split3")
    Nn = B3.shape[0]
    out = torch.empty((M, Nn), dtype=torch.float32, device=A.device)
    L.check(lib.cnx_gemm_plain(L.ptr(a3), L.ptr(B3), L.ptr(bias), L.ptr(out), L.dt(torch.float32), M, Nn, 3 * K, L.dt(bf),
                               L.CNX_GEMM_A_SPLIT2 if seg2 else 0, L.stream()), "gemm_plain(x3)")
    return out


def _ln_fwd(x2, w, b, eps, out_dtype):
    lib = L.load()
    M, C = x2.shape
    out = torch.empty((M, C), dtype=out_dtype, device=x2.device)
    mean = torch.empty((M,), dtype=torch.float32, device=x2.device)
    rstd = torch.empty((M,), dtype=torch.float32, device=x2.device)
    L.check(lib.cnx_ln_fwd(L.ptr(x2), L.dt(x2), L.ptr(w), L.ptr(b), eps, M, C, L.ptr(out), L.dt(out_dtype), L.ptr(mean),
                           L.ptr(rstd), L.stream()), "ln_fwd")
    return out, mean, rstd


def _ln_bwd(dxn, y, mean, rstd, w, dy_dtype, params=None):
    """-> dy [M,C], d ln_w [C], d ln_b [C] (None, None when `params` = (ln_w, ln_b) leaves have a gradient sink)"""
    lib = L.load()
    M, C = y.shape
    P = max(1, min(_num_partials(C), (M + 7) // 8))
    dy = torch.empty((M, C), dtype=dy_dtype, device=y.device)
    part = torch.empty((P, 2 * C), dtype=torch.float32, device=y.device)
    L.check(lib.cnx_ln_bwd(L.ptr(dxn), L.dt(dxn), L.ptr(y), L.dt(y), L.ptr(mean), L.ptr(rstd), L.ptr(w), M, C, L.ptr(dy),
                           L.dt(dy_dtype), L.ptr(part), P, L.stream()), "ln_bwd")
    d = _Dest(params if params is not None else (w, w), ((C,), (C,))) if params is not None else None
    if d is None:
        dwb = torch.empty((2 * C,), dtype=torch.float32, device=y.device)
        L.check(lib.cnx_reduce_partials(L.ptr(part), P, 2 * C, 1.0, 0, L.ptr(dwb), L.stream()), "reduce_partials")
        return dy, dwb[:C], dwb[C:]
    L.check(lib.cnx_reduce_partials_split(L.ptr(part), P, C, C, d.acc, L.ptr(d.bufs[0]), L.ptr(d.bufs[1]), L.stream()),
            "reduce_partials_split")
    dlw, dlb = d.results()
    return dy, dlw, dlb


def _patch_weight(conv_w: torch.Tensor, act_dtype, channels_last_taps: bool) -> torch.Tensor:
    """[Cout,Cin,k,k] conv weight -> the GEMM's B operand [Cout, k*k*Cin] in the activation dtype.  The stem operand
    keeps the canonical (ci,ky,kx) flattening; the downsample operand is (ky,kx,ci) to match channels-last patches."""
    def build():
        Cout = conv_w.shape[0]
        w2 = conv_w.detach().permute(0, 2, 3, 1) if channels_last_taps else conv_w.detach()
        return w2.reshape(Cout, -1).to(act_dtype).contiguous()
    return _derived((conv_w,), ("patchw", channels_last_taps, act_dtype), build)


class _StemFn(torch.autograd.Function):
    """stem: Conv2d(3, C, 4, stride 4) -> LayerNorm2d (convnext.py:79-82) as patchify + tcgen05 GEMM + LayerNorm kernels.
    Under bf16 autocast the conv output is rounded to bf16 and the LayerNorm output is fp32, as ATen's policies give."""

    @staticmethod
    def forward(ctx, x, conv_w, conv_b, ln_w, ln_b, eps, act_dtype, track):
        lib = L.load()
        L.require_cuda(x, conv_w, ln_w)
        N, Cin, H, W = x.shape
        Cout = conv_w.shape[0]
        xc = x.detach().to(torch.float32).contiguous()
        M = N * (H // 4) * (W // 4)
        A = torch.empty((M, Cin * 16), dtype=act_dtype, device=x.device)
        L.check(lib.cnx_patchify4_nchw(L.ptr(xc), N, Cin, H, W, L.ptr(A), L.dt(act_dtype), L.stream()), "patchify4")
        if not track and X3_FWD and act_dtype == torch.float32 and Cout % 8 == 0:
            y = _gemm_plain_x3(A, conv_w, False, conv_b)
        else:
            y = _gemm_plain(A, _patch_weight(conv_w, act_dtype, False), conv_b, act_dtype)
        out, mean, rstd = _ln_fwd(y, ln_w, ln_b, eps, torch.float32)
        if track and any(ctx.needs_input_grad[1:5]):
            ctx.save_for_backward(A, y, mean, rstd, conv_w, ln_w)
            ctx.act_dtype = act_dtype
            ctx.params = (conv_w, conv_b, ln_w, ln_b)
        return out.view(N, H // 4, W // 4, Cout).permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dout):
        A, y, mean, rstd, conv_w, ln_w = ctx.saved_tensors
        p_conv_w, p_conv_b, p_ln_w, p_ln_b = ctx.params
        M, Cout = y.shape
        d2 = _nhwc(dout).reshape(M, Cout)
        dy, dlw, dlb = _ln_bwd(d2, y, mean, rstd, ln_w, ctx.act_dtype, params=(p_ln_w, p_ln_b))
        d = _Dest((p_conv_w, p_conv_b), (conv_w.shape, (Cout,))) if p_conv_w.is_contiguous() else None
        if d is None:
            dW, db = _wgrad(dy, A, M, Cout, A.shape[1], True)
            return None, dW.view_as(conv_w), db, dlw, dlb, None, None, None
        _wgrad(dy, A, M, Cout, A.shape[1], True, out=d.bufs[0], cs=d.bufs[1], accumulate=d.acc)   # [Cout, Cin*16] IS the conv weight's layout
        dW, db = d.results()
        return None, dW, db, dlw, dlb, None, None, None


class _DownsampleFn(torch.autograd.Function):
    """downsample: LayerNorm2d -> Conv2d(C, C2, 2, stride 2) (convnext.py:84-89) as LayerNorm + 2x2 patch gather + GEMM."""

    @staticmethod
    def forward(ctx, x, ln_w, ln_b, conv_w, conv_b, eps, act_dtype, track, widen):
        global _LAST_OUT
        _LAST_OUT = None
        lib = L.load()
        L.require_cuda(x, conv_w, ln_w)
        N, C, H, W = x.shape
        C2 = conv_w.shape[0]
        xl = _nhwc(x.detach())
        M = N * H * W
        # `widen`: the consumer (a Block with an fp32 layer scale, modules.ConvNeXtStage) widens this bf16 conv output to fp32 on
        # entry (ATen type promotion in the reference, convnext.py:55): the GEMM then writes the fp32 tensor itself, holding the
        # bf16-rounded values — no bf16 tensor, no cast pass; in backward the Block hands back the bf16 copy of its dx
        widen = bool(widen) and act_dtype == torch.bfloat16 and GEMM_FLAGS == 0 and C2 % 8 == 0
        # LayerNorm writes the GEMM operand directly in 2x2-patch-major order (no gather pass)
        A = torch.empty((M // 4, 4 * C), dtype=act_dtype, device=x.device)
        mean = torch.empty((M,), dtype=torch.float32, device=x.device)
        rstd = torch.empty((M,), dtype=torch.float32, device=x.device)
        L.check(lib.cnx_ln_fwd_patch2(L.ptr(xl), L.dt(xl), L.ptr(ln_w), L.ptr(ln_b), eps, N, H, W, C, L.ptr(A), L.dt(act_dtype),
                                      L.ptr(mean), L.ptr(rstd), L.stream()), "ln_fwd_patch2")
        if not track and X3_FWD and act_dtype == torch.float32 and C2 % 8 == 0:
            out = _gemm_plain_x3(A, conv_w, True, conv_b)
        elif widen:
            out = _gemm_plain(A, _patch_weight(conv_w, act_dtype, True), conv_b, torch.float32, round_bf16=True)
        else:
            out = _gemm_plain(A, _patch_weight(conv_w, act_dtype, True), conv_b, act_dtype)
        if track and (ctx.needs_input_grad[0] or any(ctx.needs_input_grad[1:5])):
            ctx.save_for_backward(xl, mean, rstd, A, conv_w, ln_w)
            ctx.shape = (N, C, H, W)
            ctx.act_dtype = act_dtype
            if widen and DZ_HANDOFF:
                _LAST_OUT = (out.data_ptr(), (N, H // 2, W // 2, C2), None)
        return out.view(N, H // 2, W // 2, C2).permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dout):
        lib = L.load()
        xl, mean, rstd, A, conv_w, ln_w = ctx.saved_tensors
        N, C, H, W = ctx.shape
        act_dtype = ctx.act_dtype
        C2 = conv_w.shape[0]
        M = N * H * W
        global _DZ_HANDOFF
        ho, _DZ_HANDOFF = _DZ_HANDOFF, None
        dl = _nhwc(dout)
        if (ho is not None and ho[2] is None and ho[0].data_ptr() == dl.data_ptr() and ho[0].shape == dl.shape
                and ho[0].dtype == dl.dtype and ho[1].dtype == act_dtype):
            d2 = ho[1].reshape(M // 4, C2)       # bf16 copy of this gradient, written by the first Block's dwconv backward-data kernel
        else:
            d2 = dl.reshape(M // 4, C2)
            if d2.dtype != act_dtype:
                d2 = d2.to(act_dtype)
        del ho
        # weight / bias gradient: dWp[C2, 4C] = d2^T . A  (taps-last layout) -> canonical [C2, C, 2, 2]
        dWp, db = _wgrad(d2, A, M // 4, C2, 4 * C, True)
        dW = dWp.view(C2, 2, 2, C).permute(0, 3, 1, 2).contiguous()
        dx = dlw = dlb = None
        # data gradient: dA = d2 . Wp  (B operand = Wp^T [4C, C2]) -> scatter back to pixel order -> LayerNorm backward
        wpt = _derived((conv_w,), ("patchwT", act_dtype),
                       lambda: conv_w.detach().permute(2, 3, 1, 0).reshape(4 * C, C2).to(act_dtype).contiguous())
        dA = _gemm_plain(d2, wpt, None, act_dtype)            # patch-major; the LayerNorm backward reads it in place
        P = max(1, min(_num_partials(C), (M + 7) // 8))
        dxl = torch.empty((M, C), dtype=xl.dtype, device=dA.device)
        part = torch.empty((P, 2 * C), dtype=torch.float32, device=dA.device)
        L.check(lib.cnx_ln_bwd_patch2(L.ptr(dA), L.dt(dA), L.ptr(xl), L.dt(xl), L.ptr(mean), L.ptr(rstd), L.ptr(ln_w), N, H, W, C,
                                      L.ptr(dxl), L.dt(xl.dtype), L.ptr(part), P, L.stream()), "ln_bwd_patch2")
        dwb = torch.empty((2 * C,), dtype=torch.float32, device=dA.device)
        L.check(lib.cnx_reduce_partials(L.ptr(part), P, 2 * C, 1.0, 0, L.ptr(dwb), L.stream()), "reduce_partials")
        dlw, dlb = dwb[:C], dwb[C:]
        dx = dxl.view(N, H, W, C).permute(0, 3, 1, 2)
        return dx, dlw, dlb, dW, db, None, None, None, None


def stem_forward(x, conv_w, conv_b, ln_w, ln_b, eps: float):
    """Conv2d(Cin, C, 4, 4) + LayerNorm2d on a [N,Cin,H,W] image batch -> logical [N,C,H/4,W/4] (channels-last memory)."""
    return _StemFn.apply(x, conv_w, conv_b, ln_w, ln_b, float(eps), _act_dtype(), torch.is_grad_enabled())


def downsample_forward(x, ln_w, ln_b, conv_w, conv_b, eps: float, widen: bool = False):
    """LayerNorm2d + Conv2d(C, C2, 2, 2) on a logical [N,C,H,W] stream -> logical [N,C2,H/2,W/2].  widen: under bf16 autocast
    return the conv output as the fp32 tensor its consumer would widen it to (same values)."""
    return _DownsampleFn.apply(x, ln_w, ln_b, conv_w, conv_b, float(eps), _act_dtype(), torch.is_grad_enabled(), bool(widen))


class _HeadFn(torch.autograd.Function):
    """timm NormMlpClassifierHead (global avg pool -> LayerNorm2d -> flatten -> fc) on libcnx kernels: cnx_avgpool_nhwc,
    cnx_ln_fwd / cnx_ln_bwd on the [N, C] rows, fc as the tcgen05 GEMM (bf16 autocast) with its wgrad + bias column sums.
    dtypes follow ATen's autocast policies: pooled and normalised features fp32, logits in the autocast dtype."""

    @staticmethod
    def forward(ctx, x, ln_w, ln_b, fc_w, fc_b, eps, act_dtype, track):
        lib = L.load()
        L.require_cuda(x, ln_w, fc_w)
        N, C, H, W = x.shape
        K = fc_w.shape[0]
        xl = _nhwc(x.detach())
        pooled = torch.empty((N, C), dtype=torch.float32, device=x.device)
        L.check(lib.cnx_avgpool_nhwc_fwd(L.ptr(xl), L.dt(xl), N, H * W, C, L.ptr(pooled), L.stream()), "avgpool_fwd")
        xn, mean, rstd = _ln_fwd(pooled, ln_w, ln_b, eps, act_dtype)      # fp32 LayerNorm, rounded once to the GEMM operand dtype
        if act_dtype == torch.float32:
            if not track and X3_FWD and C % 8 == 0:
                logits = _gemm_plain_x3(xn, fc_w, False, fc_b)
            else:
                logits = _gemm_plain(xn, fc_w, fc_b, torch.float32)
        else:
            logits = _gemm_plain(xn, _weight_prep(fc_w, 0, None, act_dtype), fc_b, act_dtype)
        if track and (ctx.needs_input_grad[0] or any(ctx.needs_input_grad[1:5])):
            ctx.save_for_backward(pooled, xn, mean, rstd, ln_w, fc_w)
            ctx.meta = (N, C, H, W, xl.dtype, act_dtype)
            ctx.params = (ln_w, ln_b, fc_w, fc_b)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        lib = L.load()
        pooled, xn, mean, rstd, ln_w, fc_w = ctx.saved_tensors
        N, C, H, W, sdt, act_dtype = ctx.meta
        K = fc_w.shape[0]
        d = dlogits.contiguous()
        if d.dtype != act_dtype:
            d = d.to(act_dtype)
        # d xn = d logits . W   (B operand = W^T [C, K]);  dW = d logits^T . xn, db = column sums of d logits
        wt = _weight_prep(fc_w, 1, None, act_dtype) if act_dtype != torch.float32 else _derived(
            (fc_w,), ("fcT",), lambda: fc_w.detach().t().contiguous())
        dxn = _gemm_plain(d, wt, None, act_dtype)
        p_ln_w, p_ln_b, p_fc_w, p_fc_b = ctx.params
        dfc = _Dest((p_fc_w, p_fc_b), ((K, C), (K,)))
        _wgrad(d, xn, N, K, C, True, out=dfc.bufs[0], cs=dfc.bufs[1], accumulate=dfc.acc)
        dW, db = dfc.results()
        dpooled, dlw, dlb = _ln_bwd(dxn, pooled, mean, rstd, ln_w, torch.float32, params=(p_ln_w, p_ln_b))
        dx = None
        if ctx.needs_input_grad[0]:
            dxl = torch.empty((N, H, W, C), dtype=sdt, device=d.device)
            L.check(lib.cnx_avgpool_nhwc_bwd(L.ptr(dpooled), N, H * W, C, L.ptr(dxl), L.dt(sdt), L.stream()), "avgpool_bwd")
            dx = dxl.permute(0, 3, 1, 2)
        return dx, dlw, dlb, dW, db, None, None, None


def head_forward(x, ln_w, ln_b, fc_w, fc_b, eps: float):
    """NormMlpClassifierHead forward on a logical [N,C,H,W] stream -> logits [N,K]."""
    return _HeadFn.apply(x, ln_w, ln_b, fc_w, fc_b, float(eps), _act_dtype(), torch.is_grad_enabled())


def head_supported(x, fc_w) -> bool:
    """shapes the kernels take: C a multiple of 8 (16-byte bf16 rows), K a multiple of 8 (GEMM N / wgrad N1)"""
    return x.is_cuda and x.dim() == 4 and x.shape[1] % 8 == 0 and fc_w.shape[0] % 8 == 0 and x.dtype in (torch.float32, torch.bfloat16)


def mixup_batch(x: torch.Tensor, lam: float, box=None, original_out: torch.Tensor | None = None) -> torch.Tensor:
    """timm Mixup._mix_batch's tensor work on a contiguous fp32 CUDA batch, in place and in one pass (bit-exact with the
    flip / mul_ / mul_ / add_ sequence): `box=(yl, yh, xl, xh)` selects the cutmix box swap.  `original_out`, when given,
    receives the un-mixed batch (the second device copy engine.py:40 makes)."""
    lib = L.load()
    L.require_cuda(x)
    if x.dtype != torch.float32 or not x.is_contiguous() or x.dim() != 4:
        raise TypeError("mixup_batch: contiguous fp32 [B,C,H,W] batch expected")
    if original_out is not None and (original_out.shape != x.shape or original_out.dtype != x.dtype
                                     or not original_out.is_contiguous() or original_out.device != x.device):
        raise TypeError("mixup_batch: original_out must match the batch")
    B, C, H, W = x.shape
    yl, yh, xl, xh = (int(v) for v in box) if box is not None else (0, 0, 0, 0)
    L.check(lib.cnx_mixup_batch(L.ptr(x), L.ptr(original_out), B, C, H, W, float(lam), int(box is not None), yl, yh, xl, xh,
                                L.stream()), "mixup_batch")
    return x


class _SoftTargetCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, target):
        lib = L.load()
        L.require_cuda(x, target)
        if x.dim() != 2 or target.shape != x.shape:
            raise ValueError(f"SoftTargetCrossEntropy expects x and target of the same [B,K] shape, got {tuple(x.shape)} "
                             f"and {tuple(target.shape)}")
        B, K = x.shape
        xc = x.detach().contiguous()
        tc = target.detach().to(torch.float32).contiguous()
        dev = x.device
        loss = torch.empty((), dtype=torch.float32, device=dev)
        lse = torch.empty((B,), dtype=torch.float32, device=dev)
        row = torch.empty((B,), dtype=torch.float32, device=dev)
        counter = torch.zeros((1,), dtype=torch.int32, device=dev)
        L.check(lib.cnx_soft_target_ce_fwd(L.ptr(xc), L.dt(xc), L.ptr(tc), B, K, L.ptr(loss), L.ptr(lse), L.ptr(row),
                                           L.ptr(counter), L.stream()), "soft_target_ce_fwd")
        ctx.save_for_backward(xc, tc, lse)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        lib = L.load()
        xc, tc, lse = ctx.saved_tensors
        B, K = xc.shape
        dl = dloss.detach().to(torch.float32).contiguous()
        dx = torch.empty_like(xc)
        L.check(lib.cnx_soft_target_ce_bwd(L.ptr(xc), L.dt(xc), L.ptr(tc), L.ptr(lse), L.ptr(dl), B, K, L.ptr(dx),
                                           L.dt(dx), L.stream()), "soft_target_ce_bwd")
        return dx, None


def soft_target_cross_entropy(x, target):
    return _SoftTargetCEFn.apply(x, target)


def mixup_target(target: torch.Tensor, num_classes: int, lam: float = 1.0, smoothing: float = 0.0) -> torch.Tensor:
    """timm.data.mixup.mixup_target on the GPU, bit-exact in fp32 (one kernel instead of six)."""
    lib = L.load()
    L.require_cuda(target)
    t = target.detach().long().contiguous().view(-1)
    B = t.numel()
    out = torch.empty((B, num_classes), dtype=torch.float32, device=t.device)
    L.check(lib.cnx_mixup_target(L.ptr(t), B, num_classes, float(lam), float(smoothing), L.ptr(out), L.stream()),
            "mixup_target")
    return out


# ------------------------------------------------------------------------------------------------------------------------
# torch.library registration: the entry points above are also torch custom ops, `torch.ops.cnx.*` (what the nn.Module mirror in
# modules.py / loss.py / mixup.py calls).  They are registered CompositeImplicitAutograd: the dispatcher hands the call to the
# Python function above, whose torch.autograd.Function records the backward — so autograd, autocast state and grad mode are
# exactly what a direct call sees.  Schemas are the contract a maintainer would bind against.
# ------------------------------------------------------------------------------------------------------------------------
_LIB = torch.library.Library("cnx", "DEF")
_SCHEMAS = {
    "block_forward": ("(Tensor x, Tensor conv_w, Tensor conv_b, Tensor ln_w, Tensor ln_b, Tensor w1, Tensor b1, Tensor w2, Tensor b2, "
                      "Tensor? gamma, Tensor? dp, float eps) -> Tensor", block_forward),
    "layer_norm_cl": ("(Tensor x, Tensor w, Tensor b, float eps) -> Tensor", layer_norm_cl),
    "stem_forward": ("(Tensor x, Tensor conv_w, Tensor conv_b, Tensor ln_w, Tensor ln_b, float eps) -> Tensor", stem_forward),
    "downsample_forward": ("(Tensor x, Tensor ln_w, Tensor ln_b, Tensor conv_w, Tensor conv_b, float eps, bool widen=False) -> Tensor",
                           downsample_forward),
    "head_forward": ("(Tensor x, Tensor ln_w, Tensor ln_b, Tensor fc_w, Tensor fc_b, float eps) -> Tensor", head_forward),
    "soft_target_cross_entropy": ("(Tensor x, Tensor target) -> Tensor", soft_target_cross_entropy),
    "mixup_target": ("(Tensor target, int num_classes, float lam=1.0, float smoothing=0.0) -> Tensor", mixup_target),
    "mixup_batch": ("(Tensor(a!) x, float lam, int[]? box=None, Tensor(b!)? original_out=None) -> Tensor(a!)", mixup_batch),
}
for _name, (_schema, _fn) in _SCHEMAS.items():
    _LIB.define(_name + _schema)
    _LIB.impl(_name, _fn, "CompositeImplicitAutograd")
