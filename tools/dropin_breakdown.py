#!/usr/bin/env python
"""Where one drop-in call spends its time (host side): gs_insert.gs_watermark_init_noise step by step, the torch route it takes
today against the host-buffer pipe (gswm_pipe_embed_injected), and extract.recover_exactracted_message likewise.
Usage: python tools/dropin_breakdown.py   (run under gpurun)"""
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
from gswm import codec, extract, gs_insert  # noqa: E402

os.chdir(tempfile.mkdtemp())
dev = torch.device("cuda:0")
key, nonce = bytes.fromhex(gswm.DEFAULT_KEY_HEX), bytes.fromhex(gswm.DEFAULT_NONCE_HEX)
k = codec.pad_message("lthero", 32)
km = codec.KeyMaterial.make(key, nonce, k, 256)
pipe = gswm.HostPipe(0, max_elems=16384, chunk_latents=1)
N = 200


def med(f, n=N):
    for _ in range(10):
        f()
    t = []
    for _ in range(n):
        t0 = time.perf_counter()
        f()
        t.append(time.perf_counter() - t0)
    return round(1e6 * float(np.median(t)), 1)


u = np.random.uniform(0, 1, size=16384)
du = torch.from_numpy(u).to(dev)
z = codec.embed_batch_injected(du, (4, 64, 64), km, 1, torch.float64)
zh = z[0].cpu().numpy()
lat16 = torch.from_numpy(zh).half().reshape(1, 4, 64, 64)
args = types.SimpleNamespace(key=key, nonce=nonce, l=1, message_length=256)
opt = types.SimpleNamespace(key_hex=gswm.DEFAULT_KEY_HEX, nonce_hex=gswm.DEFAULT_NONCE_HEX)
out = {
    "embed": {
        "np.random.uniform(16384) [the reference's own draw]": med(lambda: np.random.uniform(0, 1, size=16384)),
        "pad_message + resolve_key_nonce + KeyMaterial.make": med(lambda: codec.KeyMaterial.make(*codec.resolve_key_nonce(opt.key_hex, opt.nonce_hex), codec.pad_message("lthero", 32), 256)),
        "torch.from_numpy(u).to(device)": med(lambda: torch.from_numpy(u).to(dev)),
        "embed_batch_injected (launch, no sync)": med(lambda: codec.embed_batch_injected(du, (4, 64, 64), km, 1, torch.float64)),
        "z[0].cpu().numpy()": med(lambda: z[0].cpu().numpy()),
        "torch route: upload + kernel + download": med(lambda: codec.embed_batch_injected(torch.from_numpy(u).to(dev), (4, 64, 64), km, 1, torch.float64)[0].cpu().numpy()),
        "pipe route: gswm_pipe_embed_injected (host in, host out)": med(lambda: pipe.embed_injected(u, (4, 64, 64), km, 1, np.float64)),
        "info_data.txt append": med(lambda: gs_insert._append_info(key, nonce, k)),
        "whole call": med(lambda: gs_insert.gs_watermark_init_noise(opt, "lthero")),
    },
    "extract": {
        "recover_exactracted_message (whole call)": med(lambda: extract.recover_exactracted_message(lat16, args)),
        "z.to(device) of the fp16 latent": med(lambda: lat16.reshape(1, -1).to(dev)),
        "extract_batch (launch) + flags.cpu()": med(lambda: codec.extract_batch(lat16.reshape(1, -1).to(dev), codec.KeyMaterial.make(key, nonce, None, 256)).flags.cpu()),
        "pipe route: gswm_pipe_extract (host in, host out)": med(lambda: pipe.extract(lat16.reshape(1, -1), codec.KeyMaterial.make(key, nonce, None, 256))),
        "calculate_bit_accuracy": med(lambda: extract.calculate_bit_accuracy((b"lthero" + bytes(26)).hex(), "0" * 256)),
    },
    "unit": "us, median of %d" % N,
}
print(json.dumps(out, indent=1))
