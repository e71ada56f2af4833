"""Drop-in for ``scripts/GS_watermark_insert_for_webui_v1.6.0_and_higher.py`` (AUTOMATIC1111 webui >= 1.6.0).

Same globals, ``init_gs_Z_s_T``, ``modified_ImageRNG`` (``first()`` returns the watermarked noise, ``next()`` falls
back to ordinary randn) and ``Script`` (v1.6.0:17-190).  Two upstream defects are not reproduced: the reference
restores ``rng.ImageRNG`` to the *modified* class in its ``finally`` (v1.6.0:190) and refers to undefined
``rng_philox`` / ``nv_rng`` (v1.6.0:103,111); here the original class is restored and the NV generator comes
from ``modules.rng_philox``.
"""
from __future__ import annotations

import torch

from . import webui_v152 as _w
from .webui_v152 import set_seed  # noqa: F401  (same helper, v1.6.0:150-153)

# the state lives in one place so both webui generations share init_gs_Z_s_T (v1.6.0:17-23)
global_use_treering = 0


def __getattr__(name):           # global_message, global_key, ... read through to the shared module state
    if name.startswith("global_"):
        return getattr(_w, name)
    raise AttributeError(name)


def init_gs_Z_s_T():
    """v1.6.0:26-91 (identical arithmetic to the <= 1.5.2 script)."""
    return _w.init_gs_Z_s_T()


def create_generator(seed):
    """v1.6.0:100-106."""
    from modules import devices, shared
    if shared.opts.randn_source == "NV":
        from modules import rng_philox
        return rng_philox.Generator(seed)
    device = devices.cpu if shared.opts.randn_source == "CPU" or devices.device.type == "mps" else devices.device
    return torch.Generator(device).manual_seed(int(seed))


def randn_without_seed(shape, generator=None):
    """v1.6.0:108-116."""
    from modules import devices, shared
    if shared.opts.randn_source == "NV":
        return torch.asarray(generator.randn(shape), device=devices.device)
    if shared.opts.randn_source == "CPU" or devices.device.type == "mps":
        return torch.randn(shape, device=devices.cpu, generator=generator).to(devices.device)
    return torch.randn(shape, device=devices.device, generator=generator)


class modified_ImageRNG:
    """v1.6.0:118-147: stands in for modules.rng.ImageRNG while the script runs."""

    def __init__(self, shape, seeds, subseeds=None, subseed_strength=0.0, seed_resize_from_h=0, seed_resize_from_w=0):
        self.shape = tuple(map(int, shape))
        self.seeds = seeds
        self.subseeds = subseeds
        self.subseed_strength = subseed_strength
        self.seed_resize_from_h = seed_resize_from_h
        self.seed_resize_from_w = seed_resize_from_w
        self.generators = [create_generator(seed) for seed in seeds]
        self.is_first = True

    def first(self):
        noise = torch.tensor(init_gs_Z_s_T()).float().to(_w._shared_device())
        return noise.unsqueeze(0)

    def next(self):
        if self.is_first:
            self.is_first = False
            return self.first()
        xs = [randn_without_seed(self.shape, generator=g) for g in self.generators]
        return torch.stack(xs).to(_w._shared_device())


def _make_script():
    import gradio as gr
    import modules.rng as rng
    import modules.scripts as scripts
    from modules.processing import process_images

    from . import codec

    class Script(scripts.Script):
        def title(self):
            return "GS_watermark_insert"

        def ui(self, is_img2img):
            key_input = gr.Textbox(label="Input Key Here", value=codec.DEFAULT_KEY_HEX)
            nonce_input = gr.Textbox(label="Input Nonce Here", value=codec.DEFAULT_NONCE_HEX)
            message_input = gr.Textbox(label="Input Message Here", value="")
            use_repeat = gr.Textbox(label="1 means repeat message four times, 0 means not", value="0")
            use_randomSeed_input = gr.Textbox(label="1 means use use_randomSeed, 0 means not", value="0")
            with gr.Row():
                seed_input = gr.Number(label="Seed", value="42")
                seed_button = gr.Button("Generate Random Seed")
            seed_button.click(fn=set_seed, inputs=None, outputs=seed_input)
            return [message_input, key_input, nonce_input, seed_input, use_randomSeed_input, use_repeat]

        def run(self, p, message, key, nonce, seed, use_randomSeed, use_repeat):
            real_rng = rng.ImageRNG
            try:
                rng.ImageRNG = modified_ImageRNG
                _w.global_message, _w.global_key, _w.global_nonce = message, key, nonce
                _w.global_randomSeed = int(set_seed(seed))
                _w.global_use_randomSeed = int(use_randomSeed)
                _w.global_use_repeat = int(use_repeat)
                return process_images(p)
            finally:
                rng.ImageRNG = real_rng          # the reference leaves the patch in place (v1.6.0:190)

    return Script


try:
    Script = _make_script()
except Exception:  # noqa: BLE001
    Script = None
