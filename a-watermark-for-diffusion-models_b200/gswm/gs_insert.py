"""Drop-in for the reference's ``gs_insert.py``: same function name, arguments, return value and side
effect (``info_data.txt``), with the arithmetic done by libgswm.so on the GPU.

    from gswm.gs_insert import gs_watermark_init_noise          # instead of `from gs_insert import ...`
    start_code = torch.stack([torch.tensor(gs_watermark_init_noise(opt, opt.message)).float()
                              for _ in range(opt.n_samples)]).to(device)            # README.md:110-112

``gs_watermark_init_noise_batch`` is the batched, device-resident form of that loop.
"""
from __future__ import annotations

from datetime import datetime

import numpy as np
import torch

from . import _embed_common as common
from . import codec


def _append_info(key: bytes, nonce: bytes, k: bytes, path: str = "info_data.txt") -> None:
    # gs_insert.py:68-74, same fields and order
    current_time = datetime.now().strftime("%Y-%m-%d %H:%M:%S")
    with open(path, "a") as f:
        f.write(f"Time: {current_time}\n")
        f.write(f"key: {key.hex()}\n")
        f.write(f"nonce: {nonce.hex()}\n")
        f.write(f"message: {k.hex()}\n")
        f.write("----------------------\n")


def gs_watermark_init_noise(opt, message=""):
    """gs_insert.py:8-75.  Returns a float64 numpy array of shape (4, 64, 64).

    ``opt`` needs ``key_hex`` and ``nonce_hex`` (README.md:52-70).  Uniforms are drawn from numpy's
    global generator exactly as the reference's per-element ``np.random.uniform(0, 1)`` calls would
    (same stream, same consumption), then injected into the float64 device path, so a seeded numpy
    state reproduces the reference's latent to ~1e-12 relative.
    """
    k = codec.pad_message(message, 32)                                   # gs_insert.py:9-20
    key, nonce = codec.resolve_key_nonce(opt.key_hex, opt.nonce_hex)     # gs_insert.py:27-42
    u = np.random.uniform(0, 1, size=4 * 64 * 64)                        # gs_insert.py:62, one per element
    z = common.embed_injected_host(u, (4, 64, 64), key, nonce, k, 256, 1, np.float64)
    _append_info(key, nonce, k)
    return z[0]


def gs_watermark_init_noise_batch(opt, message="", n_samples=1, seed=None, device=None, first_latent=0,
                                  log=True) -> torch.Tensor:
    """The README.md:110-112 loop as one launch: (n_samples, 4, 64, 64) fp32 on ``device``.

    Every sample gets its own uniforms from the in-kernel counter-based generator (``seed`` defaults
    to fresh OS entropy, like the reference's unseeded numpy state).  One ``info_data.txt`` record
    is appended, because key / nonce / message are shared by the batch.
    """
    import os

    k = codec.pad_message(message, 32)
    key, nonce = codec.resolve_key_nonce(opt.key_hex, opt.nonce_hex)
    if seed is None:
        seed = int.from_bytes(os.urandom(8), "little")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    km = codec.KeyMaterial.make(key, nonce, k, 256)
    out = codec.embed_batch(n_samples, (4, 64, 64), km, seed, 0, first_latent, dev)
    if log:
        _append_info(key, nonce, k)
    return out
