"""Shared host logic of the reference-named embed shims (gs_insert / webui / ComfyUI): draw the
uniforms exactly where the reference draws them, run the float64 device path, append info_data.txt."""
from __future__ import annotations

from datetime import datetime
from typing import Optional, Sequence

import numpy as np
import torch

from . import codec


def draw_uniforms(n: int, seeded: bool, seed: Optional[int], copies: int = 1) -> np.ndarray:
    """The reference's per-element ``u = np.random.uniform(0, 1)`` / ``rng.uniform(0, 1)`` draws
    (gs_insert.py:62; nodes.py:52-53,114-117; v1.5.2:27,72-75), vectorised: the legacy MT19937 stream
    yields the same values whether drawn one at a time or as an array."""
    if seeded:
        return np.random.RandomState(seed=seed).uniform(0, 1, size=n)
    return np.random.uniform(0, 1, size=(copies, n)) if copies > 1 else np.random.uniform(0, 1, size=n)


def seed_is_mt_int(seed) -> bool:
    """True when ``RandomState(seed)`` takes numpy's integer path (init_genrand): the case the device generator restates."""
    return isinstance(seed, (int, np.integer)) and not isinstance(seed, bool) and 0 <= int(seed) <= 0xFFFFFFFF


def embed_seeded(seed, latent_shape: Sequence[int], key: bytes, nonce: bytes, k: bytes, msg_bits: int,
                 out_dtype: torch.dtype, device=None) -> torch.Tensor:
    """One latent whose uniforms are ``RandomState(seed).uniform(0, 1)`` per element (nodes.py:52-53,117; v1.5.2:27,75).
    For an integer seed the whole thing runs on the GPU -- MT19937 stream, Phi^-1, scatter -- and nothing is uploaded;
    any other seed numpy accepts (None, arrays) draws on the host like the reference and takes the injected path.
    Returns a device tensor [1, *latent_shape]."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if seed_is_mt_int(seed):
        km = codec.KeyMaterial.make(key, nonce, k, msg_bits)
        return codec.embed_batch_mt19937(int(seed), 1, latent_shape, km, out_dtype, dev)
    n = int(np.prod(latent_shape))
    return embed_injected(draw_uniforms(n, True, seed), latent_shape, key, nonce, k, msg_bits, 1, out_dtype, dev)


def embed_injected(u: np.ndarray, latent_shape: Sequence[int], key: bytes, nonce: bytes, k: bytes, msg_bits: int,
                   n_latents: int, out_dtype: torch.dtype, device=None) -> torch.Tensor:
    """z = norm.ppf((u + y) / 2) on the GPU in float64 for n_latents latents; returns a device tensor."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    km = codec.KeyMaterial.make(key, nonce, k, msg_bits)
    return codec.embed_batch_injected(torch.from_numpy(np.ascontiguousarray(u)).to(dev), latent_shape, km, n_latents, out_dtype)


_pipes = {}                      # device index -> (HostPipe, max_elems); the reference's call sites are single-threaded, a lock keeps
_pipes_lock = __import__("threading").Lock()     # concurrent callers from interleaving on one pipe anyway


def host_pipe(n_elems: int, device_index=None) -> "codec.HostPipe":
    """The process's host-buffer pipe for ``device_index`` (created on first use, regrown for larger latents): the drop-ins
    take host arrays in and hand host arrays back, and gswm_pipe_* does that with one pinned staging copy each way and no
    torch tensor in between -- 55 us instead of 130 us per single-latent call (tools/dropin_breakdown.py)."""
    idx = torch.cuda.current_device() if device_index is None else int(device_index)
    cur = _pipes.get(idx)
    if cur is None or cur[1] < n_elems:
        if cur is not None:
            cur[0].close()
        cap = max(65536, int(n_elems))
        # a chunk holds up to ~64 MB of float64 per buffer
        cur = (codec.HostPipe(idx, max_elems=cap, chunk_latents=max(1, min(64, (8 << 20) // cap))), cap)
        _pipes[idx] = cur
    return cur[0]


def embed_injected_host(u: np.ndarray, latent_shape: Sequence[int], key: bytes, nonce: bytes, k: bytes, msg_bits: int,
                        n_latents: int, out_dtype=np.float64) -> np.ndarray:
    """z = norm.ppf((u + y) / 2) on the GPU in float64 for host uniforms, result back on the host: numpy
    [n_latents, *latent_shape] of ``out_dtype`` (float64 as gs_insert.py:75 returns it, or float32)."""
    km = codec.KeyMaterial.make(key, nonce, k, msg_bits)
    n = int(np.prod(latent_shape))
    with _pipes_lock:
        return host_pipe(n).embed_injected(u, latent_shape, km, n_latents, out_dtype)


def append_info(lines: Sequence[str], path: str = "info_data.txt") -> None:
    """Append one record to ./info_data.txt the way every reference variant does (gs_insert.py:68-74)."""
    current_time = datetime.now().strftime("%Y-%m-%d %H:%M:%S")
    with open(path, "a") as f:
        f.write(f"Time: {current_time}\n")
        for ln in lines:
            f.write(ln + "\n")
        f.write("----------------------\n")
