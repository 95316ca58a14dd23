"""train_one_epoch — the reference's step engine (engine.py:10-143) with the same signature, argument meaning and
return value, for callers that want the step loop from this package instead of the reference's file.  The
reference's own `engine.train_one_epoch` also works unmodified with this package's model / criterion / EMA / mixup
objects (that is the drop-in boundary; INTEGRATION.md shows the two-line change in train.py).

What is kept exactly (results identical to engine.py on the same inputs):
  * lr / weight-decay schedule writes (engine.py:33-38), H2D copies (:40-41), mixup (:43-44), forward + loss (:46-52),
    `loss.item()` + non-finite guard that skips the step (:54-59), `loss /= update_freq; backward; step every
    update_freq; zero_grad; model_ema.update` (:61-77), the second no-grad forward on the un-mixed batch for train
    accuracy when mixup is on (:89-97), per-class TP/FP/FN totals and the returned {loss, class_acc} global averages.
Placement of the accuracy forward: the reference runs `model(original_samples)` (:89-97) OUTSIDE its autocast block, i.e. in
fp32 even when `use_amp=True`; so does this function by default (`acc_forward_fp32=True`: fp32-accurate split-operand tcgen05
GEMMs, include/cnx.h "x3").  `acc_forward_fp32=False` runs it under the same bf16 autocast as the training forward instead —
cheaper (bench.py reports both), parameters / EMA / loss unaffected, `class_acc` and the TP/FP/FN counts can differ on argmax
near-ties.
What differs, without changing results:
  * `use_amp=True` means bf16 autocast (the BASELINE north-star precision; the reference's torch.amp.autocast('cuda')
    defaults to fp16).  A `loss_scaler` object is driven exactly as engine.py:61-68 drives it (it owns
    backward, unscale, clip / grad-norm and optimizer.step); bf16 needs no loss scaling, so this package's own
    `utils.NativeScalerWithGradNormCount` keeps utils.py:427-453's interface without a GradScaler.  With `loss_scaler=None`
    the step is backward -> (clip when max_norm) -> optimizer.step.  Without AMP nothing clips, as in engine.py:70-77.
  * per-class TP/FP/FN are accumulated ON THE DEVICE with three index_add_ calls per step instead of 3*num_classes
    `.item()` host syncs (engine.py:84-87/93-96), and read back once at the end of the epoch.
  * host batches reach the device one step AHEAD: the H2D copy of batch i+1 (engine.py:40-41's `.to(device, non_blocking=True)`)
    is issued on a side stream while step i computes (`DevicePrefetcher`); same tensors, same values, no PCIe time on the
    compute stream.  `prefetch=False` restores the in-line copy.
  * for the duration of an epoch the cyclic garbage collector is tuned (`tune_gc=True`: live objects frozen, young-generation
    threshold raised; restored on return) — it cost ~1 ms of a 35 ms step.
  * rich progress bar / tensorboard / wandb plumbing is the caller's business: `log_writer` / `wandb_logger` are
    accepted and fed the same keys, but nothing is imported here.
"""
from __future__ import annotations

import gc
import math
import time
from typing import Iterable, Optional

import torch

from .mixup import Mixup as _Mixup
from .utils import clip_grad_norm_


def _add_class_counts(tp, pc, tc, preds, targets):
    """tp / pc / tc (int64 device vectors [num_classes]) += true positives / predicted count / target count per class, with NO
    host synchronisation: `torch.bincount` reads max(input) back to size its output and a boolean-mask index reads the hit
    count back — either one drains the stream at the end of every step (measured: the launching thread lost its whole lead,
    0.5 ms of idle GPU per step).  index_add_ on int64 counters is atomic integer addition: exact and order-independent."""
    one = torch.ones_like(preds)
    tp.index_add_(0, preds, (preds == targets).to(torch.int64))
    pc.index_add_(0, preds, one)
    tc.index_add_(0, targets, one)


_SCALAR_SLOTS: dict = {}         # device -> (pinned fp32 ring, next index)


def _async_scalar(t: torch.Tensor, device):
    """Enqueue the device->host copy of a 0-dim tensor NOW (4 bytes into page-locked memory, event recorded behind it) and return
    a function that waits for that copy alone — not for whatever is enqueued afterwards — and returns the Python float."""
    ring = _SCALAR_SLOTS.get(device)
    if ring is None:
        ring = _SCALAR_SLOTS[device] = [torch.empty(8, dtype=torch.float32).pin_memory(), 0]
    buf, i = ring
    ring[1] = (i + 1) % buf.numel()
    slot = buf[i:i + 1]
    slot.copy_(t.detach().reshape(1), non_blocking=True)
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(device))

    def wait() -> float:
        ev.synchronize()
        return float(slot[0])
    return wait


class DevicePrefetcher:
    """Iterate a loader of (samples, targets) one batch ahead: the host->device copies of the NEXT batch are issued on a
    side stream before the current batch is handed out, so they overlap the current step's kernels.  Batches that are
    already on the device pass through untouched.

    The copies land in two device buffers per tensor shape; the side stream and the buffers are created once per device and
    kept for the life of the process (no stream creation, no caching-allocator traffic and no cudaMalloc inside an epoch).
    Ordering: the consumer's stream waits on the copy's event before it touches a batch; before a buffer is overwritten
    (two batches later, or by the next epoch) the side stream waits on an event recorded on the consumer's stream when the
    consumer asked for the following batch, i.e. after all work on the buffer's previous batch was enqueued.  A consumer
    that keeps a batch beyond the next TWO `next()` calls must clone it."""

    _state = {}                    # device index -> {"stream", "bufs", "released": [event | None, event | None]}

    def __init__(self, loader, device):
        self.loader = loader
        self.device = torch.device(device)

    def __len__(self):
        return len(self.loader)

    def _dev_state(self):
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        st = DevicePrefetcher._state.get(idx)
        if st is None:
            st = DevicePrefetcher._state[idx] = {"stream": torch.cuda.Stream(self.device), "bufs": {}, "released": [None, None],
                                                 "busy": False}
        if st["busy"]:
            # a second loader iterated while another one is live on this device (nested loops): private stream and buffers
            return {"stream": torch.cuda.Stream(self.device), "bufs": {}, "released": [None, None], "busy": True, "private": True}
        st["busy"] = True
        return st

    def __iter__(self):
        dev = self.device
        it = iter(self.loader)
        st = None
        count = 0

        def fetch():
            nonlocal st, count
            try:
                s, t = next(it)
            except StopIteration:
                return None
            if s.is_cuda and t.is_cuda:
                return s, t, None, None
            if st is None:
                st = self._dev_state()
            side, bufs, released = st["stream"], st["bufs"], st["released"]
            slot = count & 1
            count += 1
            out = []
            with torch.cuda.stream(side):
                if released[slot] is not None:
                    side.wait_event(released[slot])
                for which, h in enumerate((s, t)):
                    if h.is_cuda:
                        out.append(h)
                        continue
                    key = (slot, which, tuple(h.shape), h.dtype)
                    d = bufs.get(key)
                    if d is None:
                        d = bufs[key] = torch.empty(h.shape, dtype=h.dtype, device=dev)
                    d.copy_(h, non_blocking=True)
                    out.append(d)
                ev = torch.cuda.Event()
                ev.record(side)
            return out[0], out[1], ev, slot

        def release(slot):
            # everything the consumer did with the batch in `slot` is on its stream by now
            e = torch.cuda.Event()
            e.record(torch.cuda.current_stream(dev))
            st["released"][slot] = e

        nxt = fetch()
        prev_slot = None
        try:
            while nxt is not None:
                s, t, ev, slot = nxt
                if prev_slot is not None:
                    release(prev_slot)
                if ev is not None:
                    torch.cuda.current_stream(dev).wait_event(ev)
                prev_slot = slot
                if st is not None and st.get("private"):
                    # private buffers die with this iterator: tell the caching allocator that the consumer's stream uses them,
                    # so that their memory is not handed out again while kernels enqueued there still read it
                    for buf in (s, t):
                        if buf.is_cuda:
                            buf.record_stream(torch.cuda.current_stream(dev))
                nxt = fetch()
                yield s, t
        finally:
            # the buffers outlive this epoch: the next one must not overwrite them under the last steps of this one
            if st is not None:
                for sl in (0, 1):
                    release(sl)
                st["busy"] = False


def train_one_epoch(model: torch.nn.Module, criterion: torch.nn.Module, data_loader: Iterable,
                    optimizer: torch.optim.Optimizer, device: torch.device, epoch: int, loss_scaler=None,
                    max_norm: float = 0, model_ema=None, mixup_fn=None, log_writer=None, wandb_logger=None,
                    start_steps: Optional[int] = 0, lr_schedule_values=None, wd_schedule_values=None,
                    num_training_steps_per_epoch: Optional[int] = None, update_freq: Optional[int] = 1,
                    use_amp: bool = False, num_classes: int = 2, verbose: bool = True,
                    prefetch: bool = True, acc_forward_fp32: bool = True, tune_gc: bool = True):
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("imageclassification_b200.engine runs on CUDA devices only (no CPU fallback); "
                           "use oracle/engine.py for CPU reference numbers")
    update_freq = update_freq or 1
    start_steps = start_steps or 0
    if num_training_steps_per_epoch is None:
        num_training_steps_per_epoch = len(data_loader) // update_freq
    model.train(True)
    optimizer.zero_grad()
    start_time = time.time()
    tp = torch.zeros(num_classes, dtype=torch.int64, device=device)
    pc = torch.zeros_like(tp)
    tc = torch.zeros_like(tp)
    acc_sum = torch.zeros((), dtype=torch.float32, device=device)

    batches = DevicePrefetcher(data_loader, device) if prefetch else data_loader
    # Python's cyclic collector was measured to cost ~1 ms of a 35 ms step here (hundreds of young-generation passes per step over
    # an object graph dominated by long-lived model / optimizer state).  For the duration of the epoch: everything alive now
    # is frozen out of the collector's reach and the young-generation threshold is raised; both are restored on return.
    gc_state = None
    if tune_gc and gc.isenabled():
        gc_state = (gc.get_threshold(), gc.get_freeze_count() == 0)
        gc.collect()
        if gc_state[1]:                      # objects the host application froze itself stay frozen: only undo our own freeze
            gc.freeze()
        gc.set_threshold(max(gc_state[0][0], 50_000), gc_state[0][1], gc_state[0][2])
    try:
        return _train_loop(model, criterion, batches, optimizer, device, loss_scaler, max_norm, model_ema, mixup_fn, log_writer,
                           wandb_logger, start_steps, lr_schedule_values, wd_schedule_values, num_training_steps_per_epoch,
                           update_freq, use_amp, num_classes, verbose, acc_forward_fp32, tp, pc, tc, acc_sum, start_time)
    finally:
        if gc_state is not None:
            gc.set_threshold(*gc_state[0])
            if gc_state[1]:
                gc.unfreeze()


def _train_loop(model, criterion, batches, optimizer, device, loss_scaler, max_norm, model_ema, mixup_fn, log_writer, wandb_logger,
                start_steps, lr_schedule_values, wd_schedule_values, num_training_steps_per_epoch, update_freq, use_amp, num_classes,
                verbose, acc_forward_fp32, tp, pc, tc, acc_sum, start_time):
    loss_total, n_updates = 0.0, 0
    for data_iter_step, (samples, targets) in enumerate(batches):
        step = data_iter_step // update_freq
        if step >= num_training_steps_per_epoch:
            continue
        it = start_steps + step
        if lr_schedule_values is not None or wd_schedule_values is not None and data_iter_step % update_freq == 0:
            for group in optimizer.param_groups:
                if lr_schedule_values is not None:
                    group["lr"] = lr_schedule_values[it]
                if wd_schedule_values is not None and group["weight_decay"] > 0:
                    group["weight_decay"] = wd_schedule_values[it]

        samples = samples.to(device, non_blocking=True)
        targets = targets.to(device, non_blocking=True)
        original_samples, original_targets = samples, targets
        if mixup_fn is not None:
            # the reference makes a second device copy (engine.py:40) because mixup mutates `samples` in place
            if isinstance(mixup_fn, _Mixup):
                original_samples = torch.empty_like(samples)         # filled by the mixing kernel itself
                samples, targets = mixup_fn(samples, targets, original_out=original_samples)
            else:
                original_samples = samples.clone() if samples.is_cuda else samples
                samples, targets = mixup_fn(samples, targets)

        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bool(use_amp)):
            output = model(samples)
            loss = criterion(output, targets)

        grad_norm = None
        update_grad = (data_iter_step + 1) % update_freq == 0
        if use_amp and loss_scaler is not None:
            # engine.py:54-68: backward, unscale, gradient norm / clipping and the optimizer step all belong to the caller's
            # scaler object (utils.py:427-447; this package's bf16 one is `NativeScalerWithGradNormCount` in utils.py), which
            # runs backward and the step in ONE call: the loss has to be on the host before it is called
            loss_value = loss.item()
            if not math.isfinite(loss_value):
                print("Loss is {}, stopping training".format(loss_value))
                optimizer.zero_grad()
                continue
            loss /= update_freq
            grad_norm = loss_scaler(loss, optimizer, clip_grad=max_norm, parameters=model.parameters(), create_graph=False,
                                    update_grad=update_grad)
            if update_grad:
                optimizer.zero_grad()
                if model_ema is not None:
                    model_ema.update(model)
        else:
            # engine.py:54-59,70-77 (no AMP: never clips) — and bf16 autocast without a scaler object, where `max_norm` clips.
            # The loss value travels to the host through a pinned word whose copy is enqueued BEFORE backward, and is read only
            # after backward has been enqueued: the launching thread never waits for an idle GPU in the middle of a step
            # (`loss.item()` at engine.py:54 drains the stream: measured 0.56 ms of idle GPU + ~1 ms of launch-bound backward per
            # 35 ms step).  Same decisions as the reference: a non-finite loss leaves parameters, optimizer state and EMA
            # untouched and discards every accumulated gradient (engine.py:56-59 calls zero_grad(), so whether backward ran
            # first cannot be observed); the step's metrics are skipped.
            pending = _async_scalar(loss, device)
            loss /= update_freq
            loss.backward()
            loss_value = pending()
            if not math.isfinite(loss_value):
                print("Loss is {}, stopping training".format(loss_value))
                optimizer.zero_grad()
                continue
            if update_grad:
                if use_amp and max_norm:
                    grad_norm = clip_grad_norm_(model.parameters(), max_norm)
                optimizer.step()
                optimizer.zero_grad()
                if model_ema is not None:
                    model_ema.update(model)

        with torch.no_grad():
            if mixup_fn is None:
                ref_out, ref_t = output, targets
            else:
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bool(use_amp) and not acc_forward_fp32):
                    ref_out = model(original_samples)
                ref_t = original_targets
            preds = ref_out.max(1)[1]
            _add_class_counts(tp, pc, tc, preds, ref_t)
            class_acc = (preds == ref_t).float().mean()
            acc_sum += class_acc

        loss_total += loss_value
        n_updates += 1
        if log_writer is not None:
            log_writer.update(loss=loss_value, head="loss")
            log_writer.update(class_acc=class_acc, head="loss")
            log_writer.update(lr=max(g["lr"] for g in optimizer.param_groups), head="opt")
            log_writer.update(min_lr=min(g["lr"] for g in optimizer.param_groups), head="opt")
            if grad_norm is not None:
                log_writer.update(grad_norm=grad_norm, head="opt")
            log_writer.set_step()
        if wandb_logger:
            wandb_logger._wandb.log({"Rank-0 Batch Wise/train_loss": loss_value,
                                     "Rank-0 Batch Wise/global_train_step": it})

    stats = {"loss": loss_total / max(n_updates, 1), "class_acc": (acc_sum / max(n_updates, 1)).item()}
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        # utils.py:80-88: global average over ranks of (count, total) per meter
        t = torch.tensor([n_updates, loss_total, stats["class_acc"] * n_updates], dtype=torch.float64, device=device)
        torch.distributed.all_reduce(t)
        n = max(t[0].item(), 1.0)
        stats = {"loss": t[1].item() / n, "class_acc": t[2].item() / n}
    tp_l, pc_l, tc_l = tp.tolist(), pc.tolist(), tc.tolist()
    stats_extra = {"true_positives": tp_l, "false_positives": [p - t for p, t in zip(pc_l, tp_l)],
                   "false_negatives": [c - t for c, t in zip(tc_l, tp_l)]}
    if verbose:
        print(f"Averaged stats:loss: {stats['loss']:.4f}  class_acc: {stats['class_acc']:.4f},Time:{time.time() - start_time}")
        for i in range(num_classes if num_classes <= 16 else 0):
            fp_i, fn_i = stats_extra["false_positives"][i], stats_extra["false_negatives"][i]
            precision = tp_l[i] / (tp_l[i] + fp_i) if tp_l[i] + fp_i > 0 else 0
            recall = tp_l[i] / (tp_l[i] + fn_i) if tp_l[i] + fn_i > 0 else 0
            print(f"Class {i}: Precision: {precision:.5f}, Recall: {recall:.5f}")
    train_one_epoch.last_class_counts = stats_extra
    return stats


@torch.no_grad()
def evaluate(data_loader, model, device, num_classes, use_amp=False, verbose: bool = True, prefetch: bool = True):
    """The reference's validation loop (engine.py:145-225): same arguments, same returned keys and values
    ({loss, acc1, avg_precision, avg_recall, precision_i, recall_i} as global averages).  The forward runs the libcnx kernels
    (no-grad path: fused MLP where the hidden activation fits on chip); per-class TP / predicted / target counts and the loss and
    top-1 sums stay on the device (three index_add_ calls per batch instead of 3*num_classes `.item()` syncs) and are read back once."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("imageclassification_b200.engine runs on CUDA devices only (no CPU fallback); "
                           "use oracle/engine.py for CPU reference numbers")
    criterion = torch.nn.CrossEntropyLoss()
    model.eval()
    tp = torch.zeros(num_classes, dtype=torch.int64, device=device)
    pc = torch.zeros_like(tp)
    tc = torch.zeros_like(tp)
    sums = torch.zeros(2, dtype=torch.float64, device=device)          # sum of batch losses, number of correct top-1
    n_batches = n_samples = 0
    batches = DevicePrefetcher(_first_last(data_loader), device) if prefetch else _first_last(data_loader)
    for images, target in batches:
        images = images.to(device, non_blocking=True)
        target = target.to(device, non_blocking=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bool(use_amp)):
            output = model(images)
            loss = criterion(output, target)
        preds = output.max(1)[1]
        _add_class_counts(tp, pc, tc, preds, target)
        sums[0] += loss.double()
        sums[1] += (preds == target).sum().double()    # top-1 of timm accuracy(): argmax, first index on ties as topk gives
        n_batches += 1
        n_samples += int(images.shape[0])
    loss_sum, correct = sums.tolist()
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        # utils.py:80-88 synchronize_between_processes: (count, total) of every meter summed over ranks
        t = torch.tensor([n_batches, loss_sum, n_samples, correct], dtype=torch.float64, device=device)
        torch.distributed.all_reduce(t)
        n_batches, loss_sum, n_samples, correct = t.tolist()
    stats = {"loss": loss_sum / max(n_batches, 1), "acc1": 100.0 * correct / max(n_samples, 1)}
    tp_l, pc_l, tc_l = tp.tolist(), pc.tolist(), tc.tolist()
    precs, recs = [], []
    for i in range(num_classes):
        precs.append(tp_l[i] / pc_l[i] if pc_l[i] > 0 else 0)          # TP + FP = predicted count
        recs.append(tp_l[i] / tc_l[i] if tc_l[i] > 0 else 0)           # TP + FN = target count
        stats[f"precision_{i}"] = precs[-1]
        stats[f"recall_{i}"] = recs[-1]
        if verbose and num_classes <= 16:
            print(f"Class {i}: Precision: {precs[-1]:.5f}, Recall: {recs[-1]:.5f}")
    stats["avg_precision"] = sum(precs) / len(precs)
    stats["avg_recall"] = sum(recs) / len(recs)
    if verbose:
        print(f"Average Precision: {stats['avg_precision']:.5f}, Average Recall: {stats['avg_recall']:.5f}")
    return stats


class _first_last:
    """engine.py:170-171 takes `batch[0]` and `batch[-1]` of whatever the loader yields."""

    def __init__(self, loader):
        self.loader = loader

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        for batch in self.loader:
            yield batch[0], batch[-1]
