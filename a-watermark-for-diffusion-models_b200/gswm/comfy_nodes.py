"""Drop-in for ``ComfyUI_GSWaterMark/nodes.py``: ``gs_watermark_init_noise`` with the reference's signature,
the ``GSLatent`` / ``GSKSamplerAdvanced`` nodes and ``NODE_CLASS_MAPPINGS`` (nodes.py:26-252).  The per-element
loop is replaced by the GPU float64 path; a whole batch of independent latents is one launch.

ComfyUI's ``comfy.*`` / ``latent_preview`` are imported only when a sampler node actually runs, so the codec
functions work (and are tested) outside ComfyUI.
"""
from __future__ import annotations

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
    u = common.draw_uniforms(n, int(use_seed) == 1, randomSeed)                     # nodes.py:52-53,114-117
    z = common.embed_injected(u, (4, height // 8, width // 8), key, nonce, k, bits, 1, torch.float32)
    _log(key, nonce, k, randomSeed, height, width, message_length)
    return z[0].cpu()


def gs_watermark_init_noise_batch(key_hex, nonce_hex, message, batch_size, width, height, message_length=-1):
    """batch_size independent unseeded latents (nodes.py:237) as one launch: (B, 4, h, w) fp32 CPU tensor.
    Uniforms come from numpy's global generator in the order the reference's sequential calls would draw them."""
    n, bits, k = _frame(message, width, height, message_length)
    key, nonce = codec.resolve_key_nonce(key_hex, nonce_hex)
    u = common.draw_uniforms(n, False, None, copies=batch_size).reshape(batch_size, n)
    z = common.embed_injected(u, (4, height // 8, width // 8), key, nonce, k, bits, batch_size, torch.float32)
    for _ in range(batch_size):
        _log(key, nonce, k, None, height, width, message_length)
    return z.cpu()


def common_ksampler(model, seed, steps, cfg, sampler_name, scheduler, positive, negative, latent, denoise=1.0,
                    disable_noise=False, start_step=None, last_step=None, force_full_denoise=False, use_GS=False,
                    GS_latent_noise=None):
    """nodes.py:141-164 (ComfyUI sampler glue; the watermarked noise replaces prepare_noise when use_GS)."""
    import comfy.sample
    import comfy.utils
    import latent_preview

    latent_image = latent["samples"]
    if use_GS:
        noise = GS_latent_noise["samples"]
    elif disable_noise:
        noise = torch.zeros(latent_image.size(), dtype=latent_image.dtype, layout=latent_image.layout, device="cpu")
    else:
        noise = comfy.sample.prepare_noise(latent_image, seed, latent.get("batch_index"))
    callback = latent_preview.prepare_callback(model, steps)
    samples = comfy.sample.sample(model, noise, steps, cfg, sampler_name, scheduler, positive, negative, latent_image,
                                  denoise=denoise, disable_noise=disable_noise, start_step=start_step, last_step=last_step,
                                  force_full_denoise=force_full_denoise, noise_mask=latent.get("noise_mask"),
                                  callback=callback, disable_pbar=not comfy.utils.PROGRESS_BAR_ENABLED, seed=seed)
    out = latent.copy()
    out["samples"] = samples
    return (out,)


def _sampler_choices():
    try:
        import comfy.samplers
        return comfy.samplers.KSampler.SAMPLERS, comfy.samplers.KSampler.SCHEDULERS
    except Exception:  # noqa: BLE001  (outside ComfyUI)
        return ["euler"], ["normal"]


class GSKSamplerAdvanced:
    """nodes.py:167-207."""

    @classmethod
    def INPUT_TYPES(s):
        samplers, schedulers = _sampler_choices()
        return {"required": {
            "model": ("MODEL",),
            "add_GS_noise": (["enable", "disable"],),
            "add_noise": (["disable", "enable"],),
            "noise_seed": ("INT", {"default": 42, "min": 0, "max": 0xffffffffffffffff}),
            "steps": ("INT", {"default": 20, "min": 1, "max": 10000}),
            "cfg": ("FLOAT", {"default": 8.0, "min": 0.0, "max": 100.0, "step": 0.1, "round": 0.01}),
            "sampler_name": (samplers,),
            "scheduler": (schedulers,),
            "positive": ("CONDITIONING",),
            "negative": ("CONDITIONING",),
            "latent_image": ("LATENT",),
            "GS_latent_noise": ("LATENT",),
            "start_at_step": ("INT", {"default": 0, "min": 0, "max": 10000}),
            "end_at_step": ("INT", {"default": 10000, "min": 0, "max": 10000}),
            "return_with_leftover_noise": (["disable", "enable"],),
        }}

    RETURN_TYPES = ("LATENT",)
    FUNCTION = "sample"
    CATEGORY = "GSWatermark-lthero/sampling"

    def sample(self, model, add_GS_noise, add_noise, noise_seed, steps, cfg, sampler_name, scheduler, positive, negative,
               latent_image, GS_latent_noise, start_at_step, end_at_step, return_with_leftover_noise, denoise=1.0):
        return common_ksampler(model, noise_seed, steps, cfg, sampler_name, scheduler, positive, negative, latent_image,
                               denoise=denoise, disable_noise=add_noise == "disable", start_step=start_at_step,
                               last_step=end_at_step, force_full_denoise=return_with_leftover_noise != "enable",
                               use_GS=add_GS_noise == "enable", GS_latent_noise=GS_latent_noise)


class GSLatent:
    """nodes.py:210-240."""

    @classmethod
    def INPUT_TYPES(s):
        return {"required": {
            "use_seed": ("INT", {"default": 1, "min": 0, "max": 1}),
            "seed": ("INT", {"default": 42, "min": 0, "max": 0xffffffff}),
            "width": ("INT", {"default": 512, "min": 64, "max": MAX_RESOLUTION, "step": 8}),
            "height": ("INT", {"default": 512, "min": 64, "max": MAX_RESOLUTION, "step": 8}),
            "key": ("STRING", {"default": codec.DEFAULT_KEY_HEX}),
            "nonce": ("STRING", {"default": codec.DEFAULT_NONCE_HEX}),
            "message": ("STRING", {"default": "lthero"}),
            "message_length": ("INT", {"default": -1, "min": 32, "max": 1024, "step": 32}),
            "batch_size": ("INT", {"default": 1, "min": 1, "max": 64}),
        }}

    RETURN_TYPES = ("LATENT", "IMAGE")
    FUNCTION = "create_gs_latents"
    CATEGORY = "GSWatermark-lthero/latent/noise"

    def create_gs_latents(self, key, nonce, message, batch_size, use_seed, seed, width, height, message_length):
        if use_seed == 1:
            # one seeded latent replicated batch_size times (nodes.py:232-235)
            one = gs_watermark_init_noise(key, nonce, "cpu", message, use_seed, seed, width=width, height=height,
                                          message_length=message_length)
            latent = torch.stack([one for _ in range(batch_size)]).float()
        else:
            latent = gs_watermark_init_noise_batch(key, nonce, message, batch_size, width, height, message_length).float()
        return ({"samples": latent}, latent[0])


NODE_CLASS_MAPPINGS = {"Lthero_GSLatent": GSLatent, "Lthero_GS_KSamplerAdvanced": GSKSamplerAdvanced}
NODE_DISPLAY_NAME_MAPPINGS = {"Lthero_GSLatent": "GS Latent Noise", "Lthero_GS_KSamplerAdvanced": "GS KSamplerAdvanced"}
