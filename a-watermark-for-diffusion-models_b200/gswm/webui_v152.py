"""Drop-in for ``scripts/GS_watermark_insert_for_webui_v1.5.2_and_lower.py`` (AUTOMATIC1111 webui <= 1.5.2).

Copy this file's import into the webui ``scripts/`` folder in place of the reference script: same module
globals, ``init_gs_Z_s_T``, ``advanced_creator`` and ``Script`` (v1.5.2:15-138); the per-element scipy loop is
replaced by the GPU float64 path.  The webui modules are imported lazily so the codec functions also work
outside a webui process (that is how the parity tests call them).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _embed_common as common
from . import codec

# v1.5.2:15-21 -- state shared between Script.run and the patched tensor creator
global_message = ""
global_key = ""
global_nonce = ""
global_use_treering = 0
global_use_randomSeed = 0
global_randomSeed = 42
global_use_repeat = 0


def init_gs_Z_s_T():
    """v1.5.2:24-89: float64 (4, 64, 64) watermarked noise from the module globals.

    use_repeat == 1: an 8-byte message repeated four times fills the 32-byte watermark (v1.5.2:29-47).
    use_randomSeed != 0: uniforms from RandomState(global_randomSeed), else from numpy's global generator.
    Appends key / nonce / randomSeed / message to ./info_data.txt (v1.5.2:80-87).
    """
    k = codec.pad_message(global_message, 32, use_repeat=int(global_use_repeat) == 1)
    key, nonce = codec.resolve_key_nonce(global_key, global_nonce)
    if global_use_randomSeed != 0:                                  # v1.5.2:27,75: rng = RandomState(global_randomSeed)
        z = common.embed_seeded(global_randomSeed, (4, 64, 64), key, nonce, k, 256, torch.float64)[0].cpu().numpy()
    else:                                                           # v1.5.2:73: numpy's global generator
        z = common.embed_injected_host(common.draw_uniforms(4 * 64 * 64, False, None), (4, 64, 64), key, nonce, k, 256, 1,
                                       np.float64)[0]
    common.append_info([f"key: {key.hex()}", f"nonce: {nonce.hex()}", f"randomSeed: {global_randomSeed}", f"message: {k.hex()}"])
    return z


def _shared_device():
    try:
        from modules import shared
        return shared.device
    except Exception:  # noqa: BLE001  (outside webui)
        return torch.device("cuda", torch.cuda.current_device())


def advanced_creator(shape, seeds, subseeds=None, subseed_strength=0.0, seed_resize_from_h=0, seed_resize_from_w=0, p=None):
    """v1.5.2:92-97: replacement for processing.create_random_tensors -- one (1, 4, 64, 64) fp32 tensor on
    shared.device, whatever shape / batch was requested (reference behaviour, kept)."""
    noise = torch.tensor(init_gs_Z_s_T()).float().to(_shared_device())
    return noise.unsqueeze(0)


def set_seed(seed=None):
    """v1.5.2:99-102."""
    if seed is None or seed == -1:
        seed = np.random.randint(0, 2 ** 32 - 1)
    return seed


def _make_script():
    import gradio as gr
    import modules.processing as processing
    import modules.scripts as scripts
    from modules.processing import process_images

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
            global global_message, global_key, global_nonce, global_randomSeed, global_use_randomSeed, global_use_repeat
            real_creator = processing.create_random_tensors
            try:
                processing.create_random_tensors = advanced_creator
                global_message, global_key, global_nonce = message, key, nonce
                global_randomSeed = int(set_seed(seed))
                global_use_randomSeed = int(use_randomSeed)
                global_use_repeat = int(use_repeat)
                return process_images(p)
            finally:
                processing.create_random_tensors = real_creator

    return Script


try:  # inside webui: expose the Script class the loader looks for
    Script = _make_script()
except Exception:  # noqa: BLE001  (modules / gradio not importable: codec functions only)
    Script = None
