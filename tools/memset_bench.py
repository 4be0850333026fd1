#!/usr/bin/env python
"""Store-bandwidth ceiling for the affinity zero fill: torch zero_() and copy_ on a 2 GiB buffer."""
import json, torch
dev = torch.device("cuda", 0)
x = torch.empty(512 * 1024 * 1024, dtype=torch.float32, device=dev)
y = torch.empty_like(x)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
ms = t(lambda: x.zero_())
print(json.dumps(dict(op="zero_", gbs=x.numel() * 4 / ms / 1e6)))
ms = t(lambda: y.copy_(x))
print(json.dumps(dict(op="copy_", gbs_rw=2 * x.numel() * 4 / ms / 1e6)))
ms = t(lambda: x.sum())
print(json.dumps(dict(op="sum(read)", gbs=x.numel() * 4 / ms / 1e6)))
