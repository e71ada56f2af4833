"""Drop-in for ``ComfyUI_GSWaterMark/nodes.py``: ``gs_watermark_init_noise`` with the reference's signature,
the ``GSLatent`` / ``GSKSamplerAdvanced`` nodes and ``NODE_CLASS_MAPPINGS`` (nodes.py:26-252).  The per-element
loop is replaced by the GPU float64 path; a whole batch of independent latents is one launch.

ComfyUI's ``comfy.*`` / ``latent_preview`` are imported only when a sampler node actually runs, so the codec
functions work (and are tested) outside ComfyUI.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _embed_common as common
from . import codec
from .codec import choose_watermark_length  # noqa: F401  (nodes.py:26-49)

MAX_RESOLUTION = 8192


def _frame(message, width, height, message_length):
    n = 4 * (width // 8) * (height // 8)                                           # nodes.py:56-58
    bits = message_length if message_length != -1 else choose_watermark_length(n)   # nodes.py:61-64
    k = codec.pad_message(message, bits // 8)                                       # nodes.py:68-76
    return n, bits, k


def _log(key, nonce, k, randomSeed, height, width, message_length):
    # nodes.py:125-136 (randomSeed is written twice there)
    common.append_info([f"key: {key.hex()}", f"nonce: {nonce.hex()}", f"message: {k.hex()}", f"randomSeed: {randomSeed}",
                        f"height: {height}", f"width: {width}", f"randomSeed: {randomSeed}",
                        f"message_length: {message_length}"])


def gs_watermark_init_noise(key_hex, nonce_hex, device, message, use_seed, randomSeed, width, height, message_length=-1):
    """nodes.py:51-138: fp32 CPU tensor (4, height/8, width/8).  ``device`` is accepted and ignored for the result
    placement exactly as in the reference (it builds on ``device`` and returns ``.cpu()``)."""
    n, bits, k = _frame(message, width, height, message_length)
    key, nonce = codec.resolve_key_nonce(key_hex, nonce_hex)                        # nodes.py:90-99
    shape = (4, height // 8, width // 8)
    if int(use_seed) == 1:                                                          # nodes.py:52-53,117: RandomState(randomSeed)
        z = common.embed_seeded(randomSeed, shape, key, nonce, k, bits, torch.float32)[0].cpu()
    else:                                                                           # nodes.py:115: numpy's global generator
        z = torch.from_numpy(common.embed_injected_host(common.draw_uniforms(n, False, None), shape, key, nonce, k, bits, 1,
                                                        np.float32)[0])
    _log(key, nonce, k, randomSeed, height, width, message_length)
    return z


def gs_watermark_init_noise_batch(key_hex, nonce_hex, message, batch_size, width, height, message_length=-1,
                                  randomSeed=None):
    """batch_size independent unseeded latents -- the list comprehension of nodes.py:236-237 -- as one launch:
    (B, 4, h, w) fp32 CPU tensor.  What the reference does batch_size times is done batch_size times here too:
    an empty ``message`` draws a fresh ``os.urandom`` message per latent (nodes.py:76) and an empty ``key_hex`` a fresh
    key and nonce per latent (nodes.py:97-98), in the reference's call order (message, key, nonce); each latent gets
    its own ``info_data.txt`` record, with the widget's ``randomSeed`` passed through as nodes.py:131,134 log it.
    Uniforms come from numpy's global generator in the order the reference's sequential calls would draw them."""
    n = 4 * (width // 8) * (height // 8)
    rows = batch_size if (not message or not key_hex) else 1      # one row serves the whole batch when nothing is random
    ks, keys, nonces = [], [], []
    for _ in range(rows):
        _, bits, k = _frame(message, width, height, message_length)
        key, nonce = codec.resolve_key_nonce(key_hex, nonce_hex)
        ks.append(k), keys.append(key), nonces.append(nonce)
    u = common.draw_uniforms(n, False, None, copies=batch_size).reshape(batch_size, n)
    z = common.embed_injected_host(u, (4, height // 8, width // 8), b"".join(keys), b"".join(nonces), b"".join(ks), bits,
                                   batch_size, np.float32)
    for i in range(batch_size):
        r = i if rows > 1 else 0
        _log(keys[r], nonces[r], ks[r], randomSeed, height, width, message_length)
    return torch.from_numpy(z)


def common_ksampler(model, seed, steps, cfg, sampler_name, scheduler, positive, negative, latent, denoise=1.0,
                    disable_noise=False, start_step=None, last_step=None, force_full_denoise=False, use_GS=False,
                    GS_latent_noise=None):
    """Sampler glue with the reference's signature (nodes.py:141-164): the watermarked noise takes the place of
    ``prepare_noise`` when ``use_GS`` is set; everything else is handed to ``comfy.sample.sample`` untouched."""
    import comfy.sample
    import comfy.utils
    import latent_preview

    x0 = latent["samples"]
    if use_GS:
        start_noise = GS_latent_noise["samples"]
    elif disable_noise:
        start_noise = torch.zeros_like(x0, device="cpu")
    else:
        start_noise = comfy.sample.prepare_noise(x0, seed, latent.get("batch_index"))
    sampler_kwargs = dict(denoise=denoise, disable_noise=disable_noise, start_step=start_step, last_step=last_step,
                          force_full_denoise=force_full_denoise, noise_mask=latent.get("noise_mask"),
                          callback=latent_preview.prepare_callback(model, steps),
                          disable_pbar=not comfy.utils.PROGRESS_BAR_ENABLED, seed=seed)
    result = dict(latent)
    result["samples"] = comfy.sample.sample(model, start_noise, steps, cfg, sampler_name, scheduler, positive, negative, x0,
                                            **sampler_kwargs)
    return (result,)


# ---- node widgets, declared as compact tables (same names, order, defaults and ranges as nodes.py:170-186,213-223) ----
def _int(default, lo, hi, step=None):
    spec = {"default": default, "min": lo, "max": hi}
    if step is not None:
        spec["step"] = step
    return ("INT", spec)


def _sockets(*names_and_types):
    return {name: (kind,) for name, kind in names_and_types}


def _toggle(first, second):
    return ([first, second],)


def _sampler_choices():
    try:
        import comfy.samplers
        return comfy.samplers.KSampler.SAMPLERS, comfy.samplers.KSampler.SCHEDULERS
    except Exception:  # noqa: BLE001  (outside ComfyUI)
        return ["euler"], ["normal"]


class GSKSamplerAdvanced:
    """KSamplerAdvanced with a second LATENT input carrying the watermarked start noise (nodes.py:167-207)."""

    RETURN_TYPES = ("LATENT",)
    FUNCTION = "sample"
    CATEGORY = "GSWatermark-lthero/sampling"

    @classmethod
    def INPUT_TYPES(cls):
        samplers, schedulers = _sampler_choices()
        w = _sockets(("model", "MODEL"))
        w["add_GS_noise"] = _toggle("enable", "disable")
        w["add_noise"] = _toggle("disable", "enable")
        w["noise_seed"] = _int(42, 0, 2 ** 64 - 1)
        w["steps"] = _int(20, 1, 10000)
        w["cfg"] = ("FLOAT", {"default": 8.0, "min": 0.0, "max": 100.0, "step": 0.1, "round": 0.01})
        w["sampler_name"] = (samplers,)
        w["scheduler"] = (schedulers,)
        w.update(_sockets(("positive", "CONDITIONING"), ("negative", "CONDITIONING"), ("latent_image", "LATENT"),
                          ("GS_latent_noise", "LATENT")))
        w["start_at_step"] = _int(0, 0, 10000)
        w["end_at_step"] = _int(10000, 0, 10000)
        w["return_with_leftover_noise"] = _toggle("disable", "enable")
        return {"required": w}

    def sample(self, model, add_GS_noise, add_noise, noise_seed, steps, cfg, sampler_name, scheduler, positive, negative,
               latent_image, GS_latent_noise, start_at_step, end_at_step, return_with_leftover_noise, denoise=1.0):
        flags = dict(use_GS=add_GS_noise == "enable", disable_noise=add_noise == "disable",
                     force_full_denoise=return_with_leftover_noise != "enable")
        return common_ksampler(model, noise_seed, steps, cfg, sampler_name, scheduler, positive, negative, latent_image,
                               denoise=denoise, start_step=start_at_step, last_step=end_at_step,
                               GS_latent_noise=GS_latent_noise, **flags)


class GSLatent:
    """Builds the watermarked LATENT batch on the GPU (nodes.py:210-240)."""

    RETURN_TYPES = ("LATENT", "IMAGE")
    FUNCTION = "create_gs_latents"
    CATEGORY = "GSWatermark-lthero/latent/noise"

    @classmethod
    def INPUT_TYPES(cls):
        text = {"key": codec.DEFAULT_KEY_HEX, "nonce": codec.DEFAULT_NONCE_HEX, "message": "lthero"}
        w = {"use_seed": _int(1, 0, 1), "seed": _int(42, 0, 2 ** 32 - 1), "width": _int(512, 64, MAX_RESOLUTION, 8), "height": _int(512, 64, MAX_RESOLUTION, 8)}
        w.update({name: ("STRING", {"default": value}) for name, value in text.items()})
        w["message_length"] = _int(-1, 32, 1024, 32)
        w["batch_size"] = _int(1, 1, 64)
        return {"required": w}

    def create_gs_latents(self, key, nonce, message, batch_size, use_seed, seed, width, height, message_length):
        if use_seed == 1:
            # the reference embeds ONE seeded latent and replicates it batch_size times (nodes.py:232-235)
            one = gs_watermark_init_noise(key, nonce, "cpu", message, use_seed, seed, width=width, height=height,
                                          message_length=message_length)
            batch = one.float().unsqueeze(0).repeat(batch_size, 1, 1, 1)
        else:
            # batch_size independent latents (nodes.py:237): one launch instead of batch_size calls
            batch = gs_watermark_init_noise_batch(key, nonce, message, batch_size, width, height, message_length,
                                                  randomSeed=seed).float()
        return ({"samples": batch}, batch[0])


NODE_CLASS_MAPPINGS = {"Lthero_GSLatent": GSLatent, "Lthero_GS_KSamplerAdvanced": GSKSamplerAdvanced}
NODE_DISPLAY_NAME_MAPPINGS = {"Lthero_GSLatent": "GS Latent Noise", "Lthero_GS_KSamplerAdvanced": "GS KSamplerAdvanced"}
