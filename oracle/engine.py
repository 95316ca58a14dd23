"""The reference's training step restated (TEST ORACLE / CPU baseline): engine.py:27-79 plus the accuracy
bookkeeping of engine.py:82-97.  Same call signature as engine.train_one_epoch (engine.py:10-15) so the parity
tests read like a call into the reference; logging (rich progress bar, MetricLogger, tensorboard/wandb,
engine.py:99-132) is reduced to a dict of running means.  Pinned against the reference's own engine.py run in
the build container (tests/golden/make_golden.py, tests/golden/engine_step.npz).

No GradScaler path: the north-star precision is bf16 autocast (no scaler needed)."""
from __future__ import annotations

import math

import torch


def train_one_epoch(model, criterion, data_loader, optimizer, device, epoch, loss_scaler=None, max_norm=0,
                    model_ema=None, mixup_fn=None, log_writer=None, wandb_logger=None, start_steps=0,
                    lr_schedule_values=None, wd_schedule_values=None, num_training_steps_per_epoch=None,
                    update_freq=1, use_amp=False, num_classes=2, amp_dtype=torch.bfloat16):
    model.train(True)
    optimizer.zero_grad()
    tp, fp, fn = [0] * num_classes, [0] * num_classes, [0] * num_classes
    loss_sum, acc_sum, count = 0.0, 0.0, 0
    if num_training_steps_per_epoch is None:
        num_training_steps_per_epoch = len(data_loader) // update_freq
    for data_iter_step, (samples, targets) in enumerate(data_loader):
        step = data_iter_step // update_freq
        if step >= num_training_steps_per_epoch:
            continue
        it = start_steps + step
        if lr_schedule_values is not None or wd_schedule_values is not None and data_iter_step % update_freq == 0:
            for group in optimizer.param_groups:                                   # engine.py:33-38
                if lr_schedule_values is not None:
                    group["lr"] = lr_schedule_values[it]
                if wd_schedule_values is not None and group["weight_decay"] > 0:
                    group["weight_decay"] = wd_schedule_values[it]
        # engine.py:40-41 — on a CPU device .to() returns self, so "original_*" alias the mixed batch
        samples, original_samples = samples.to(device, non_blocking=True), samples.to(device, non_blocking=True)
        targets, original_targets = targets.to(device, non_blocking=True), targets.to(device, non_blocking=True)
        if mixup_fn is not None:
            samples, targets = mixup_fn(samples, targets)                          # engine.py:43-44
        if use_amp:
            with torch.amp.autocast(torch.device(device).type, dtype=amp_dtype):
                output = model(samples)
                loss = criterion(output, targets)
        else:
            output = model(samples)                                                # engine.py:51-52
            loss = criterion(output, targets)
        loss_value = loss.item()                                                   # engine.py:54
        if not math.isfinite(loss_value):                                          # engine.py:56-59
            optimizer.zero_grad()
            continue
        loss /= update_freq                                                        # engine.py:71-77
        loss.backward()
        if (data_iter_step + 1) % update_freq == 0:
            if max_norm:
                torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)
            optimizer.step()
            optimizer.zero_grad()
            if model_ema is not None:
                model_ema.update(model)
        if torch.device(device).type == "cuda":
            torch.cuda.synchronize()                                               # engine.py:79
        with torch.no_grad():                                                      # engine.py:82-97
            if mixup_fn is None:
                ref_out, ref_t = output, targets
            else:
                ref_out, ref_t = model(original_samples), original_targets
            preds = ref_out.max(1)[1]
            for i in range(num_classes):
                tp[i] += torch.sum((preds == i) & (ref_t == i)).item()
                fp[i] += torch.sum((preds == i) & (ref_t != i)).item()
                fn[i] += torch.sum((preds != i) & (ref_t == i)).item()
            class_acc = (preds == ref_t).float().mean().item()
        loss_sum += loss_value
        acc_sum += class_acc
        count += 1
    return {"loss": loss_sum / max(count, 1), "class_acc": acc_sum / max(count, 1),
            "true_positives": tp, "false_positives": fp, "false_negatives": fn}


@torch.no_grad()
def evaluate(data_loader, model, device, num_classes, use_amp=False, amp_dtype=torch.bfloat16):
    """engine.py:145-225 restated: eval-mode forward, CrossEntropyLoss, per-class TP/FP/FN by 3*K masked sums, top-1 accuracy in
    percent (timm.utils.accuracy), returned as the MetricLogger's global averages: loss = mean over batches, acc1 = mean
    weighted by batch size, precision_i / recall_i and their unweighted class averages."""
    criterion = torch.nn.CrossEntropyLoss()
    tp, fp, fn = [0] * num_classes, [0] * num_classes, [0] * num_classes
    model.eval()
    loss_sum, n_batches, acc_sum, n_samples = 0.0, 0, 0.0, 0
    for batch in data_loader:
        images, target = batch[0].to(device, non_blocking=True), batch[-1].to(device, non_blocking=True)   # engine.py:170-174
        if use_amp:
            with torch.amp.autocast(torch.device(device).type, dtype=amp_dtype):
                output = model(images)
                loss = criterion(output, target)
        else:
            output = model(images)
            loss = criterion(output, target)
        _, preds = torch.max(output, 1)
        for i in range(num_classes):                                                                       # engine.py:188-191
            tp[i] += torch.sum((preds == i) & (target == i)).item()
            fp[i] += torch.sum((preds == i) & (target != i)).item()
            fn[i] += torch.sum((preds != i) & (target == i)).item()
        maxk = min(5, output.size(1))                                                                      # timm accuracy(topk=(1, 5))
        pred = output.topk(maxk, 1, True, True)[1].t()
        correct = pred.eq(target.reshape(1, -1).expand_as(pred))
        acc1 = correct[:1].reshape(-1).float().sum(0) * 100.0 / target.size(0)
        loss_sum += loss.item()
        n_batches += 1
        acc_sum += acc1.item() * images.shape[0]
        n_samples += images.shape[0]
    out = {"loss": loss_sum / max(n_batches, 1), "acc1": acc_sum / max(n_samples, 1)}
    precs, recs = [], []
    for i in range(num_classes):                                                                           # engine.py:203-214
        precs.append(tp[i] / (tp[i] + fp[i]) if tp[i] + fp[i] > 0 else 0)
        recs.append(tp[i] / (tp[i] + fn[i]) if tp[i] + fn[i] > 0 else 0)
        out[f"precision_{i}"] = precs[-1]
        out[f"recall_{i}"] = recs[-1]
    out["avg_precision"] = sum(precs) / len(precs)
    out["avg_recall"] = sum(recs) / len(recs)
    return out
