"""DistributedDataParallel — the data-parallel wrapper of train.py:218-222 (`DistributedDataParallel(model,
device_ids=[gpu], find_unused_parameters=False)`), re-designed around a flat gradient arena.

Semantics kept from torch DDP as the reference uses it: parameters and buffers are
broadcast from rank 0 at construction; every backward all-reduces (mean over ranks) ALL parameter gradients in fp32,
in ~`bucket_cap_mb` buckets formed in reverse registration order, overlapped with the rest of backward; there is no
`no_sync` (the reference all-reduces on every micro-step of `--update_freq`); `.module` is the wrapped model.

Design: one contiguous fp32 arena holds every gradient; `param.grad` is a view into it, so a bucket is a slice of
the arena and the collective needs no flatten/unflatten copies (NCCL all-reduces the slice in place over
NVLink/NVSwitch; with NVLS the reduction happens in the switch).

How gradients reach the arena:
  * the libcnx autograd Functions (ops.py: Block, stem, head) find a `_GradSink` on their parameters and have their weight-
    gradient kernels write — or, under gradient accumulation, accumulate — STRAIGHT into the parameter's arena slot (the
    kernels' `out` / `accumulate` arguments); they return no gradient for those parameters, so autograd allocates nothing
    and copies nothing, and they report the parameter ready themselves;
  * every other parameter goes through a post-accumulate-grad hook: if `zero_grad()` had set the gradient to None, autograd
    hands the hook a fresh tensor, which is copied into the arena (one multi-tensor copy per bucket) and `param.grad`
    re-pointed.
When a bucket's last parameter is ready its `all_reduce(AVG)` is launched asynchronously (NCCL's own stream, overlapped with
the rest of backward); an autograd-engine callback queued by the first arrival waits for all buckets at the end of backward.
`overlap=False` (or CNX_DDP_OVERLAP=0) instead issues ONE all-reduce of the whole arena at the end of backward: on
NVLink 5 / NVSwitch the exposed transfer is a few hundred microseconds, and no NCCL CTA competes for SMs with the persistent
one-CTA-per-SM compute kernels of backward.

On a CPU process group (gloo; used by the CPU tests) AVG is not available, so SUM + in-place divide is used.
"""
from __future__ import annotations

import os
import weakref

import torch
import torch.distributed as dist
import torch.nn as nn


from . import ops


def _padded(numel: int) -> int:
    """arena slots start on 16-byte boundaries (the kernels' 128-bit paths); the padding words stay zero"""
    return (numel + 3) // 4 * 4


class _GradSink:
    """What ops.py finds for a parameter (ops.register_grad_sink): where to write its gradient, and whom to tell."""

    def __init__(self, ddp):
        self._ddp = weakref.ref(ddp)

    def claim(self, p):
        """-> (arena view shaped like p, accumulate flag) or None when this backward must go through autograd's own
        accumulation (reducer gone, or `p.grad` is a tensor that is not this parameter's arena slot)."""
        ddp = self._ddp()
        if ddp is None or not ddp._sinks_enabled:
            return None
        view = ddp._slot[p][1]
        g = p.grad
        if g is None:
            return view, 0
        if g.data_ptr() == view.data_ptr():
            return view, 1
        return None

    def ready(self, p):
        ddp = self._ddp()
        if p.grad is None:
            p.grad = ddp._slot[p][1]
        ddp._ready(p)


class _Bucket:
    __slots__ = ("params", "offset", "numel", "pending", "handle", "flat", "stale")

    def __init__(self):
        self.params, self.offset, self.numel = [], 0, 0
        self.pending, self.handle, self.flat, self.stale = 0, None, None, []


class DistributedDataParallel(nn.Module):
    def __init__(self, module: nn.Module, device_ids=None, output_device=None, find_unused_parameters: bool = False,
                 bucket_cap_mb: float = 25.0, process_group=None, broadcast_buffers: bool = True, overlap=None,
                 direct_grads: bool = True):
        super().__init__()
        if overlap is None:
            overlap = os.environ.get("CNX_DDP_OVERLAP", "1") != "0"
        self.overlap = bool(overlap)
        self._sinks_enabled = bool(direct_grads) and os.environ.get("CNX_DDP_DIRECT", "1") != "0"
        if find_unused_parameters:
            raise NotImplementedError("find_unused_parameters=True is not supported (the reference passes False, "
                                      "train.py:220: every parameter receives a gradient every step)")
        if not dist.is_initialized():
            raise RuntimeError("DistributedDataParallel needs an initialised torch.distributed process group")
        self.module = module
        self.process_group = process_group if process_group is not None else dist.group.WORLD
        self.world_size = dist.get_world_size(self.process_group)
        self._avg_native = dist.get_backend(self.process_group) == "nccl"
        params = [p for p in module.parameters() if p.requires_grad]
        if not params:
            raise ValueError("DistributedDataParallel: module has no trainable parameters")
        dev, dt = params[0].device, params[0].dtype
        for p in params:
            if p.device != dev or p.dtype != dt:
                raise TypeError("DistributedDataParallel: all parameters must share one device and dtype")
        # rank 0's parameters and buffers win (train.py:219-221 relies on DDP for this)
        self._broadcast([p.data for p in module.parameters()] + ([b.data for b in module.buffers()] if broadcast_buffers else []))
        # buckets in reverse registration order (gradients arrive roughly back to front); first bucket small so the
        # first collective starts early, as torch's reducer does (1 MB)
        cap = int(bucket_cap_mb * 1024 * 1024) // params[0].element_size()
        first_cap = min(cap, (1024 * 1024) // params[0].element_size())
        self.buckets: list[_Bucket] = []
        cur, off = _Bucket(), 0
        for p in reversed(params):
            limit = first_cap if not self.buckets else cap
            if cur.params and cur.numel + _padded(p.numel()) > limit:
                self.buckets.append(cur)
                cur = _Bucket()
                cur.offset = off
            cur.params.append(p)
            cur.numel += _padded(p.numel())
            off += _padded(p.numel())
        self.buckets.append(cur)
        self.arena = torch.zeros(off, dtype=dt, device=dev)
        self._slot = {}
        sink = _GradSink(self)
        for bi, b in enumerate(self.buckets):
            b.flat = self.arena[b.offset:b.offset + b.numel]
            o = b.offset
            for p in b.params:
                self._slot[p] = (bi, self.arena[o:o + p.numel()].view_as(p))
                o += _padded(p.numel())
                p.register_post_accumulate_grad_hook(self._hook)
                ops.register_grad_sink(p, sink)
        self._armed = False
        self._seen = set()
        self.comm_bytes_per_step = self.arena.numel() * self.arena.element_size()

    def _broadcast(self, tensors):
        for t in tensors:
            dist.broadcast(t, src=dist.get_global_rank(self.process_group, 0) if self.process_group is not dist.group.WORLD else 0,
                           group=self.process_group)

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    # ---- backward-time machinery -----------------------------------------------------------------
    def _arm(self):
        for b in self.buckets:
            b.pending, b.handle, b.stale = len(b.params), None, []
        self._seen = set()
        self._armed = True
        torch.autograd.Variable._execution_engine.queue_callback(self._finalize)

    def _hook(self, p):
        if not self._armed:
            self._arm()                      # (resets the buckets' stale lists: must precede the append below)
        bi, view = self._slot[p]
        if p.grad.data_ptr() != view.data_ptr():
            self.buckets[bi].stale.append((view, p.grad))
            p.grad = view
        self._ready(p)

    def _ready(self, p):
        """`p.grad` is this backward's gradient (in the arena, or queued for the copy into it): count its bucket down."""
        if not self._armed:
            self._arm()
        if id(p) in self._seen:              # a sunk parameter whose (undefined-gradient) accumulation node still ran its hook
            return
        self._seen.add(id(p))
        b = self.buckets[self._slot[p][0]]
        b.pending -= 1
        if b.pending == 0:
            if b.stale:
                torch._foreach_copy_([v for v, _ in b.stale], [g for _, g in b.stale])
                b.stale = []
            if self.world_size > 1 and self.overlap:
                op = dist.ReduceOp.AVG if self._avg_native else dist.ReduceOp.SUM
                b.handle = dist.all_reduce(b.flat, op=op, group=self.process_group, async_op=True)

    def _finalize(self):
        self._armed = False
        for b in self.buckets:
            if b.pending != 0:
                raise RuntimeError("DistributedDataParallel: a parameter received no gradient in this backward "
                                   "(find_unused_parameters is not supported)")
        if self.world_size > 1 and not self.overlap:
            op = dist.ReduceOp.AVG if self._avg_native else dist.ReduceOp.SUM
            dist.all_reduce(self.arena, op=op, group=self.process_group)
            if not self._avg_native:
                self.arena.div_(self.world_size)
            return
        for b in self.buckets:
            if b.handle is not None:
                b.handle.wait()
                b.handle = None
                if not self._avg_native:
                    b.flat.div_(self.world_size)
