#!/usr/bin/env python
"""Generate tests/golden/* by EXECUTING the unmodified reference from /root/reference.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference has no tests or golden vectors of its own, so these
fixtures are what pins the oracle (oracle/gs_oracle.py) and, through it, the CUDA path.
Every output below comes from calling the reference's own functions:

* gs_insert.gs_watermark_init_noise            (gs_insert.py:8-75)
* ComfyUI_GSWaterMark.nodes.gs_watermark_init_noise / GSLatent  (nodes.py:51-138, 210-240)
* scripts/GS_watermark_insert_for_webui_v1.5.2_and_lower.init_gs_Z_s_T (v1.5.2:24-89)
* extract.recover_exactracted_message / calculate_bit_accuracy (extract.py:72-110)
* cryptography's ChaCha20 exactly as the reference calls it (gs_insert.py:45-47)

Third-party modules the reference imports but the codec never touches (diffusers,
matplotlib, comfy.*, modules.*, gradio) are stubbed in sys.modules.  Uniform injection:
numpy.random.uniform is monkey-patched to replay a recorded RandomState stream, or the
reference's own seeded RandomState path is used.
"""
import hashlib
import importlib.util
import json
import os
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
KEY_HEX = "5822ff9cce6772f714192f43863f6bad1bf54b78326973897e6b66c3186b77a7"
NONCE_HEX = "05072fd1c2265f6f2e2a4080a2bfbdd8"


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_stubs():
    class _Any:
        SAMPLERS = ["euler"]
        SCHEDULERS = ["normal"]

        def __init__(self, *a, **k):
            pass

    _stub("diffusers", DPMSolverMultistepScheduler=_Any, StableDiffusionPipeline=_Any,
          DDIMInverseScheduler=_Any, AutoencoderKL=_Any, DPMSolverMultistepInverseScheduler=_Any)
    _stub("diffusers.utils", load_image=lambda *a, **k: None)
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            mpl = _stub("matplotlib")
            mpl.pyplot = _stub("matplotlib.pyplot")
    try:
        import torchvision  # noqa: F401
    except Exception:
        tv = _stub("torchvision")
        tv.transforms = _stub("torchvision.transforms")
    try:
        import PIL.Image  # noqa: F401
    except Exception:
        pil = _stub("PIL")
        pil.Image = _stub("PIL.Image")
    comfy = _stub("comfy")
    for sub in ("model_management", "sample", "sampler_helpers", "diffusers_load", "samplers", "sd", "utils"):
        setattr(comfy, sub, _stub("comfy." + sub))
    sys.modules["comfy.samplers"].KSampler = _Any
    _stub("latent_preview")
    modules = _stub("modules")
    modules.scripts = _stub("modules.scripts", Script=object)
    modules.processing = _stub("modules.processing", process_images=lambda p: p, slerp=None,
                               create_random_tensors=None)
    modules.devices = _stub("modules.devices")
    modules.shared = _stub("modules.shared", device="cpu")
    modules.rng = _stub("modules.rng", ImageRNG=object)
    _stub("gradio")


def _load(path, name):
    spec = importlib.util.spec_from_loader(name, importlib.machinery.SourceFileLoader(name, path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _Replay:
    """Replays a fixed float64 stream through numpy.random.uniform(0, 1)."""

    def __init__(self, u):
        self.u = np.asarray(u, dtype=np.float64)
        self.i = 0

    def __call__(self, lo=0.0, hi=1.0, size=None):
        assert size is None and lo == 0 and hi == 1
        v = self.u[self.i]
        self.i += 1
        return v


class _FakeUrandom:
    """Deterministic stand-in for os.urandom: call i returns the first n bytes of sha256-counter-mode stream i
    (tests/test_gpu_parity.py installs the same patch around the drop-in)."""

    def __init__(self):
        self.calls = 0

    def __call__(self, n):
        out, block = b"", 0
        while len(out) < n:
            out += hashlib.sha256(b"gswm-golden-urandom %d %d" % (self.calls, block)).digest()
            block += 1
        self.calls += 1
        return out[:n]


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def packed_signs(z) -> np.ndarray:
    return np.packbits((np.asarray(z, dtype=np.float64).reshape(-1) >= 0).astype(np.uint8))


def main():
    _install_stubs()
    sys.path.insert(0, REF)
    gs_insert = _load(os.path.join(REF, "gs_insert.py"), "ref_gs_insert")
    extract = _load(os.path.join(REF, "extract.py"), "ref_extract")
    nodes = _load(os.path.join(REF, "ComfyUI_GSWaterMark", "nodes.py"), "ref_nodes")
    webui = _load(os.path.join(REF, "scripts", "GS_watermark_insert_for_webui_v1.5.2_and_lower.py"), "ref_webui152")

    from cryptography.hazmat.backends import default_backend
    from cryptography.hazmat.primitives.ciphers import Cipher, algorithms

    def ks(key: bytes, nonce: bytes, n: int) -> bytes:
        enc = Cipher(algorithms.ChaCha20(key, nonce), mode=None, backend=default_backend()).encryptor()
        return enc.update(bytes(n)) + enc.finalize()

    out = {"meta": {"numpy": np.__version__}}
    arrays = {}
    import scipy
    import cryptography
    out["meta"].update(scipy=scipy.__version__, cryptography=cryptography.__version__)

    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    os.chdir(tmp)  # the reference appends ./info_data.txt
    try:
        # ---------------- ChaCha20 through the reference's library -----------------
        key = bytes.fromhex(KEY_HEX)
        cc = []
        cases = [
            ("rfc7539_2.3.2", bytes(range(32)), bytes.fromhex("01000000" "000000090000004a00000000"), 64),
            ("default_2048", key, bytes.fromhex(NONCE_HEX), 2048),
            ("default_8192", key, bytes.fromhex(NONCE_HEX), 8192),
            ("fallback_2048", key, bytes.fromhex(KEY_HEX[16:48]), 2048),
            ("fallback_8192", key, bytes.fromhex(KEY_HEX[16:48]), 8192),
            ("carry32", key, bytes.fromhex("feffffff000000000102030405060708"), 40 * 64),
            ("carry64", key, bytes.fromhex("feffffffffffffff0102030405060708"), 40 * 64),
        ]
        rs = np.random.RandomState(7)
        for i in range(8):
            cases.append((f"random{i}", rs.bytes(32), rs.bytes(16), 64 * int(rs.randint(1, 70))))
        for name, k_, n_, nb in cases:
            s = ks(k_, n_, nb)
            cc.append({"name": name, "key": k_.hex(), "nonce": n_.hex(), "nbytes": nb,
                       "first64": s[:64].hex(), "last64": s[-64:].hex(), "sha256": hashlib.sha256(s).hexdigest()})
        out["chacha20"] = cc

        # ---------------- gs_insert.gs_watermark_init_noise (SD-CLI embed) -------------
        emb = []
        opt = types.SimpleNamespace(key_hex=KEY_HEX, nonce_hex=NONCE_HEX)
        for name, message, nonce_hex, seed in [
            ("cli_lthero", "lthero", NONCE_HEX, 1234),
            ("cli_long_message", "a much longer message than thirty-two bytes, truncated", NONCE_HEX, 1235),
            ("cli_nonce_fallback", "lthero", "", 1236),
            ("cli_utf8", "水印-λ-✓", NONCE_HEX, 1237),
        ]:
            opt.nonce_hex = nonce_hex
            u = np.random.RandomState(seed).uniform(size=16384)
            real = np.random.uniform
            np.random.uniform = _Replay(u)
            try:
                z = gs_insert.gs_watermark_init_noise(opt, message)
            finally:
                np.random.uniform = real
            assert z.shape == (4, 64, 64) and z.dtype == np.float64
            arrays[name + "_z64_head"] = z.reshape(-1)[:512].copy()
            emb.append({"name": name, "message": message, "key_hex": KEY_HEX, "nonce_hex": nonce_hex,
                        "u_seed": seed, "shape": [4, 64, 64], "l_bits": 256,
                        "sha256_f64": sha(z), "sha256_f32": sha(z.astype(np.float32)),
                        "sha256_signs": sha(packed_signs(z)), "signs_head": packed_signs(z)[:8].tobytes().hex()})
        # keep one full latent (fp32, 64 KB) for value-level parity of the CUDA path
        arrays["cli_lthero_z32"] = None  # filled below
        opt.nonce_hex = NONCE_HEX
        u = np.random.RandomState(1234).uniform(size=16384)
        real = np.random.uniform
        np.random.uniform = _Replay(u)
        try:
            z_full = gs_insert.gs_watermark_init_noise(opt, "lthero")
        finally:
            np.random.uniform = real
        arrays["cli_lthero_z32"] = z_full.astype(np.float32)
        out["embed_cli"] = emb
        with open("info_data.txt") as f:
            out["info_data_cli_tail"] = f.read().splitlines()[-5:]

        # ---------------- nodes.gs_watermark_init_noise (ComfyUI, seeded) ---------------
        ne = []
        for name, seed, w, h, L, message in [
            ("comfy_512_L256", 42, 512, 512, 256, "lthero"),
            ("comfy_512_auto", 42, 512, 512, -1, "lthero"),
            ("comfy_1024_L256", 42, 1024, 1024, 256, "lthero"),
            ("comfy_1024_auto", 42, 1024, 1024, -1, "lthero"),
            ("comfy_512x1024_L256", 58, 512, 1024, 256, "lthero"),
            ("comfy_256_L32", 3, 256, 256, 32, "ab"),
            ("comfy_768x512_L96", 5, 768, 512, 96, "ninety-six!!"),
            ("comfy_64_auto", 9, 64, 64, -1, "tiny"),
        ]:
            z = nodes.gs_watermark_init_noise(KEY_HEX, NONCE_HEX, "cpu", message, 1, seed, w, h, L)
            zz = z.numpy()
            assert zz.dtype == np.float32 and zz.shape == (4, h // 8, w // 8)
            arrays[name + "_z32_head"] = zz.reshape(-1)[:256].copy()
            ne.append({"name": name, "seed": seed, "width": w, "height": h, "message_length": L, "message": message,
                       "shape": list(zz.shape), "sha256_f32": sha(zz), "sha256_signs": sha(packed_signs(zz))})
        out["embed_comfy"] = ne
        with open("info_data.txt") as f:
            out["info_data_comfy_tail"] = f.read().splitlines()[-10:]
        # GSLatent node, seeded: B identical copies (nodes.py:232-235)
        lat, first = nodes.GSLatent().create_gs_latents(KEY_HEX, NONCE_HEX, "lthero", 3, 1, 42, 512, 512, 256)
        out["gslatent_seeded"] = {"shape": list(lat["samples"].shape), "sha256": sha(lat["samples"].numpy()),
                                  "first_sha256": sha(first.numpy())}
        # GSLatent node, UNSEEDED: batch_size independent calls (nodes.py:236-237), uniforms from numpy's global
        # generator (seeded here so the run is reproducible); the widget's seed is still passed through and logged.
        np.random.seed(2024)
        lat, first = nodes.GSLatent().create_gs_latents(KEY_HEX, NONCE_HEX, "lthero", 3, 0, 77, 512, 512, 256)
        s = lat["samples"].numpy()
        arrays["gslatent_unseeded_heads"] = s.reshape(3, -1)[:, :128].copy()
        with open("info_data.txt") as f:
            tail = f.read().splitlines()[-30:]
        out["gslatent_unseeded"] = {"np_seed": 2024, "batch_size": 3, "widget_seed": 77, "shape": list(s.shape),
                                    "sha256": sha(s), "first_sha256": sha(first.numpy()),
                                    "sha256_signs_each": [sha(packed_signs(s[i])) for i in range(3)],
                                    "info_data_tail": tail}
        # ... with an EMPTY message and an EMPTY key: every one of the batch_size calls draws its own os.urandom message,
        # key and nonce (nodes.py:76,97-98).  os.urandom is replaced by a counter-mode sha256 stream so that the run can be
        # repeated by the drop-in's test (same patch, same call order: message, key, nonce per latent).
        real_urandom = os.urandom
        os.urandom = _FakeUrandom()
        try:
            np.random.seed(2025)
            lat, first = nodes.GSLatent().create_gs_latents("", "", "", 2, 0, 5, 256, 256, -1)
        finally:
            os.urandom = real_urandom
        s = lat["samples"].numpy()
        arrays["gslatent_unseeded_random_heads"] = s.reshape(2, -1)[:, :128].copy()
        with open("info_data.txt") as f:
            tail = f.read().splitlines()[-20:]
        out["gslatent_unseeded_random"] = {"np_seed": 2025, "batch_size": 2, "widget_seed": 5, "shape": list(s.shape),
                                           "sha256": sha(s), "sha256_signs_each": [sha(packed_signs(s[i])) for i in range(2)],
                                           "info_data_tail": tail}

        # ---------------- webui <=1.5.2 init_gs_Z_s_T (use_repeat + seeded) -------------
        we = []
        for name, message, use_repeat, seed in [("webui_plain", "lthero", 0, 42), ("webui_repeat", "lthero12", 1, 77),
                                                ("webui_repeat_short", "ab", 1, 5)]:
            webui.global_message = message
            webui.global_key = KEY_HEX
            webui.global_nonce = NONCE_HEX
            webui.global_use_randomSeed = 1
            webui.global_randomSeed = seed
            webui.global_use_repeat = use_repeat
            z = webui.init_gs_Z_s_T()
            arrays[name + "_z64_head"] = z.reshape(-1)[:256].copy()
            we.append({"name": name, "message": message, "use_repeat": use_repeat, "seed": seed,
                       "sha256_f64": sha(z), "sha256_f32": sha(z.astype(np.float32)),
                       "sha256_signs": sha(packed_signs(z))})
        out["embed_webui"] = we
        with open("info_data.txt") as f:
            out["info_data_webui_tail"] = f.read().splitlines()[-6:]

        # ---------------- extract.recover_exactracted_message -------------------------
        import torch
        ex = []
        base = z_full.astype(np.float32)
        msg_hex = (b"lthero" + bytes(26)).hex()
        for name, sigma, nseed, dtype in [("ext_clean_f32", 0.0, 0, "float32"), ("ext_clean_f16", 0.0, 0, "float16"),
                                          ("ext_s0325_f32", 0.325, 99, "float32"), ("ext_s1_f16", 1.0, 100, "float16"),
                                          ("ext_s43_f32", 4.3, 101, "float32"), ("ext_s8_f32", 8.0, 102, "float32"),
                                          ("ext_s20_f64", 20.0, 103, "float64")]:
            zn = base.astype(np.float64)
            if sigma:
                zn = zn + sigma * np.random.RandomState(nseed).standard_normal(zn.shape)
            zn = np.clip(zn, -60000.0, 8.0)  # stay inside what the reference can parse (extract.py:86)
            t = torch.from_numpy(zn.astype(dtype)).reshape(1, 4, 64, 64)
            args = types.SimpleNamespace(key=key, nonce=bytes.fromhex(NONCE_HEX), l=1, message_length=256)
            got = extract.recover_exactracted_message(t, args)
            orig, acc = extract.calculate_bit_accuracy(msg_hex, got)
            ex.append({"name": name, "sigma": sigma, "noise_seed": nseed, "dtype": dtype, "message_length": 256,
                       "extracted_bin": got, "bit_accuracy": acc, "original_bin": orig})
        # other L on the comfy latents
        for name, seed, w, h, L, message in [("ext_comfy_1024_L1024", 42, 1024, 1024, -1, "lthero"),
                                             ("ext_comfy_256_L32", 3, 256, 256, 32, "ab")]:
            z = nodes.gs_watermark_init_noise(KEY_HEX, NONCE_HEX, "cpu", message, 1, seed, w, h, L)
            Lb = L if L != -1 else nodes.choose_watermark_length(4 * (w // 8) * (h // 8))
            zn = z.numpy().astype(np.float64) + 1.5 * np.random.RandomState(seed).standard_normal(z.shape)
            t = torch.from_numpy(zn.astype(np.float16)).unsqueeze(0)
            args = types.SimpleNamespace(key=key, nonce=bytes.fromhex(NONCE_HEX), l=1, message_length=Lb)
            got = extract.recover_exactracted_message(t, args)
            mh = (message.encode() + bytes(Lb // 8))[:Lb // 8].hex()
            orig, acc = extract.calculate_bit_accuracy(mh, got)
            ex.append({"name": name, "seed": seed, "width": w, "height": h, "message_length": Lb, "message": message,
                       "sigma": 1.5, "dtype": "float16", "extracted_bin": got, "bit_accuracy": acc})
        out["extract"] = ex

        # quantiser edges (extract.py:83) -- one-copy latents are enough: L = N = 8
        edges = []
        for zval in [0.0, -0.0, -6e-17, 1e-300, -1e-300, -7e-17, -1e-16, -6.957291061679417e-17,
                     -6.957291061679418e-17, 5e-324, -5e-324, 1e-45, -1e-45, 8.0, -40.0, float("-inf")]:
            from scipy.stats import norm
            edges.append({"z": repr(float(zval)), "bit": int(norm.cdf(np.float64(zval)) * 2)})
        out["quantise_edges"] = edges
        raising = []
        for zval in [8.292361075813597, 9.0, float("inf"), float("nan")]:
            t = torch.full((1, 1, 2, 4), 0.5, dtype=torch.float64)
            t[0, 0, 0, 0] = zval
            args = types.SimpleNamespace(key=key, nonce=bytes.fromhex(NONCE_HEX), l=1, message_length=8)
            try:
                extract.recover_exactracted_message(t, args)
                raising.append({"z": repr(float(zval)), "raises": None})
            except Exception as e:  # noqa: BLE001
                raising.append({"z": repr(float(zval)), "raises": type(e).__name__})
        out["quantise_raises"] = raising

        # calculate_bit_accuracy corner cases (extract.py:103-110)
        ba = []
        for mh, eb in [("ff00", "1111111100000000"), ("ff00", "11111111"), ("0f", "0000111100001111"),
                       ("6c746865726f", "0" * 48), ("00" * 32, "1" * 256)]:
            o, a = extract.calculate_bit_accuracy(mh, eb)
            ba.append({"original_message_hex": mh, "extracted": eb, "original_bin": o, "accuracy": a})
        out["bit_accuracy"] = ba

        # Phi^-1 edges (gs_insert.py:64)
        from scipy.stats import norm
        out["ppf_edges"] = [{"p": repr(p), "z": repr(float(norm.ppf(p)))}
                            for p in [0.0, 0.5, 1 - 2.0 ** -53, 2.0 ** -54, 0.25, 0.75, 2.0 ** -25, 1 - 2.0 ** -25]]
    finally:
        os.chdir(cwd)

    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "golden_arrays.npz"), **arrays)
    print("wrote golden.json and golden_arrays.npz:",
          {k: (len(v) if isinstance(v, list) else 1) for k, v in out.items()})


if __name__ == "__main__":
    main()
