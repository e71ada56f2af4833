#!/usr/bin/env python
"""PCIe ceilings on the box for the e2e path: pinned H2D alone, D2H alone, and both at once (two streams)."""
import json
import time

import torch

n = 268435456
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


for _ in range(2):
    run(True, True)
a, b, c = run(True, False), run(False, True), run(True, True)
print(json.dumps({"bytes": n, "h2d_GBps": n / a / 1e9, "d2h_GBps": n / b / 1e9, "both_each_GBps": n / c / 1e9,
                  "both_ms": c * 1e3, "h2d_ms": a * 1e3, "d2h_ms": b * 1e3}))
