#!/usr/bin/env python
"""Pure-write, pure-read and copy bandwidth of this GPU with plain ATen kernels (context for the write-heavy GEMM epilogues):
CUDA events, 1 GiB buffers (larger than L2), 5 repeats, best of."""
import json
import torch

n = 1 << 30
a = torch.empty(n, dtype=torch.uint8, device="cuda")
b = torch.empty(n, dtype=torch.uint8, device="cuda")
af = a.view(torch.float32)


def t(fn, bytes_):
    for _ in range(2):
        fn()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return round(bytes_ / best / 1e6, 1)


out = {"write_zero_GBps": t(lambda: a.zero_(), n), "write_fill_f32_GBps": t(lambda: af.fill_(1.5), n),
       "copy_GBps_rw": t(lambda: b.copy_(a), 2 * n), "read_sum_f32_GBps": t(lambda: af.sum(), n),
       "read_max_u8_GBps": t(lambda: a.max(), n)}
print(json.dumps(out))
