#!/usr/bin/env python
"""BASELINE config[0] through the drop-in call sites: gs_insert.gs_watermark_init_noise(opt, 'lthero') for one SD-2.1 latent
followed by extract.recover_exactracted_message + calculate_bit_accuracy on that latent -- the call the reference takes
~2.6 s per pair for on one core.  Prints median / p90 wall-clock latency per call (host sync included).
Usage: python tools/dropin_latency.py   (run under gpurun)"""
import json
import os
import sys
import tempfile
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "a-watermark-for-diffusion-models_b200"))
import gswm  # noqa: E402
from gswm import extract, gs_insert  # noqa: E402

os.chdir(tempfile.mkdtemp())                     # info_data.txt is appended in the working directory, as in the reference
opt = types.SimpleNamespace(key_hex=gswm.DEFAULT_KEY_HEX, nonce_hex=gswm.DEFAULT_NONCE_HEX)
args = types.SimpleNamespace(key=bytes.fromhex(gswm.DEFAULT_KEY_HEX), nonce=bytes.fromhex(gswm.DEFAULT_NONCE_HEX), l=1,
                             message_length=256)
hex_msg = (b"lthero" + bytes(26)).hex()
te, tx = [], []
for i in range(60):
    t0 = time.perf_counter()
    z = gs_insert.gs_watermark_init_noise(opt, "lthero")                     # float64 numpy (4, 64, 64), result on the host
    t1 = time.perf_counter()
    latents = torch.from_numpy(z).half().reshape(1, 4, 64, 64)              # what extract.py:48,70 hands over
    t2 = time.perf_counter()
    bits = extract.recover_exactracted_message(latents, args)
    _, acc = extract.calculate_bit_accuracy(hex_msg, bits)
    t3 = time.perf_counter()
    assert acc == 1.0
    if i >= 10:
        te.append(t1 - t0)
        tx.append(t3 - t2)
q = lambda a: {"median_ms": round(1e3 * float(np.median(a)), 3), "p90_ms": round(1e3 * float(np.percentile(a, 90)), 3)}
print(json.dumps({"config": "BASELINE configs[0]: one SD-2.1 latent, default key/nonce, message 'lthero', drop-in call sites",
                  "gs_watermark_init_noise": q(te), "recover_exactracted_message+calculate_bit_accuracy": q(tx),
                  "pair_per_s": round(1.0 / (float(np.median(te)) + float(np.median(tx))), 1), "calls": len(te)}))
