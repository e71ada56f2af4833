"""gswm -- B200-native Gaussian-Shading watermark codec (embed / extract hot path).

Host side of libgswm.so (include/gswm.h).  Mirrors of the reference's call sites:

* :mod:`gswm.gs_insert`  -- ``gs_watermark_init_noise(opt, message)``           (reference gs_insert.py)
* :mod:`gswm.extract`    -- ``recover_exactracted_message``, ``calculate_bit_accuracy``  (reference extract.py)
* :mod:`gswm.codec`      -- batch API on device tensors and host buffers

There is no CPU fallback: importing works without a GPU (so the C-ABI export check can run), but
every compute call needs a CUDA device and the built library.
"""
from . import _lib, sharding
from ._lib import GswmError, build, launch_count
from .codec import (DEFAULT_KEY_HEX, DEFAULT_NONCE_HEX, Comm, ExtractResult, HostPipe, KeyMaterial, chacha20_keystream,
                    choose_watermark_length, embed_batch, embed_batch_injected, embed_batch_mt19937, embed_extract_batch,
                    extract_batch, mt19937_uniform, pad_message, resolve_key_nonce)

__all__ = ["GswmError", "build", "launch_count", "DEFAULT_KEY_HEX", "DEFAULT_NONCE_HEX", "Comm", "ExtractResult", "HostPipe",
           "KeyMaterial", "chacha20_keystream", "choose_watermark_length", "embed_batch", "embed_batch_injected",
           "embed_batch_mt19937", "embed_extract_batch", "extract_batch", "mt19937_uniform", "pad_message",
           "resolve_key_nonce", "sharding", "_lib"]
