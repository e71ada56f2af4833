#!/usr/bin/env python
"""Launch ONE kernel variant of libgswm a few times so that ncu can capture it (tools/ncu_all.sh drives this):
    python tools/ncu_targets.py {embed_shared|embed_per_latent|embed_injected|mt19937|extract_f32|extract_f16|extract_bf16|
                                 extract_per_latent|keystream} [n_latents]
B = 4096 SD-2.1 latents (4x64x64, 256-bit message) unless given; inputs are embedded + perturbed latents (sigma 0.325)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "a-watermark-for-diffusion-models_b200"))
import gswm  # noqa: E402


def main():
    which = sys.argv[1]
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    shape, n, L = (4, 64, 64), 16384, 256
    dev = torch.device("cuda:0")
    shared = gswm.KeyMaterial.make(bytes.fromhex(gswm.DEFAULT_KEY_HEX), bytes.fromhex(gswm.DEFAULT_NONCE_HEX),
                                   gswm.pad_message("lthero", L // 8), L)
    rs = np.random.RandomState(2025)
    per = gswm.KeyMaterial.make(rs.bytes(32 * B), rs.bytes(16 * B), rs.bytes((L // 8) * B), L)
    km = per if "per_latent" in which else shared
    reps = 3
    if which.startswith("embed_injected"):
        u = torch.rand((B, n), dtype=torch.float64, device=dev)
        for _ in range(reps):
            gswm.embed_batch_injected(u, shape, km)
    elif which.startswith("embed"):
        for _ in range(reps):
            gswm.embed_batch(B, shape, km, 0x5EED, 0, 0, dev)
    elif which.startswith("extract"):
        z = gswm.embed_batch(B, shape, km, 0x5EED, 0, 0, dev)
        z = z + 0.325 * torch.randn_like(z)
        z = z.to({"f16": torch.float16, "bf16": torch.bfloat16}.get(which.split("_")[1], torch.float32))
        for _ in range(reps):
            res = gswm.extract_batch(z, km)
        assert res.bit_accuracy() == 1.0
    elif which == "mt19937":
        for _ in range(reps):
            gswm.embed_batch_mt19937(list(range(1000, 1000 + B)), B, shape, shared, torch.float32, dev)
    elif which == "keystream":
        for _ in range(reps):
            gswm.chacha20_keystream(per.keys.tobytes(), per.nonces.tobytes(), 2048, dev)
    else:
        raise SystemExit(f"unknown target {which}")
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
