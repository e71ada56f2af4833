#!/usr/bin/env python
"""Does the embed kernel's two-state timing (profiles/r02_burst_states.txt) follow the SM clock?  Alternates a burst of embed
launches with the in-kernel clock measurement of gswm_debug_issue_rate (clock64 ticks / elapsed time) and prints the pairs.
Usage: python tools/clock_vs_embed.py   (run under gpurun)"""
import ctypes as C
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "a-watermark-for-diffusion-models_b200"))
import gswm  # noqa: E402
from gswm.codec import _DeviceJob  # noqa: E402

dev = torch.device("cuda:0")
lib = gswm._lib.lib()
km = gswm.KeyMaterial.make(bytes.fromhex(gswm.DEFAULT_KEY_HEX), bytes.fromhex(gswm.DEFAULT_NONCE_HEX), gswm.pad_message("lthero", 32), 256)
B, n = 4096, 16384
dj = _DeviceJob(km, B, n, dev)
z = torch.empty((B, n), dtype=torch.float32, device=dev)
st = torch.cuda.current_stream(dev).cuda_stream
rows = []
for i in range(40):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(100):
        lib.gswm_embed(C.byref(dj.job), 0x5EED, 0, 0, z.data_ptr(), st)
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / 100
    r, g = C.c_double(), C.c_double()
    lib.gswm_debug_issue_rate(0, C.byref(r), C.byref(g))          # FFMA2 chains: ~0.3 ms, reports the SM clock it ran at
    rows.append({"embed_us": round(us, 2), "sm_ghz_after": round(g.value, 3), "embed_us_x_ghz": round(us * g.value, 1)})
    time.sleep(0.05 if i % 2 else 0.3)
print(json.dumps(rows))
