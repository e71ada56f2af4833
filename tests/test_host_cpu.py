"""CPU-only checks of the host side: the C-ABI library builds/loads and exports everything
include/gswm.h declares (no compute calls without a GPU), framing / key-resolution logic against the
oracle, argument validation, and the world_size-2 sharding + counter all-reduce logic over gloo."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import gs_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gswm():
    import gswm as g
    g.build()            # nvcc cross-compiles for sm_100a without a GPU
    return g


def test_library_exports_every_declared_symbol(gswm):
    header = open(os.path.join(ROOT, "include", "gswm.h")).read()
    declared = set(re.findall(r"\b(gswm_[a-z0-9_]+)\s*\(", header))
    assert {"gswm_embed", "gswm_extract", "gswm_chacha20_keystream", "gswm_pipe_extract"} <= declared
    lib = gswm._lib.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f"libgswm.so does not export {name}"
    assert set(gswm._lib.EXPORTS) == declared
    out = subprocess.run(["nm", "-D", "--defined-only", gswm._lib.LIB_PATH], capture_output=True, text=True).stdout
    for name in declared:
        assert re.search(rf"\bT {name}\b", out), name


def test_library_is_sm100a(gswm):
    out = subprocess.run(["cuobjdump", "-lelf", gswm._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_abi_basics_without_gpu(gswm):
    lib = gswm._lib.lib()
    assert lib.gswm_abi_version() == 2 == gswm._lib.ABI_VERSION
    assert gswm._lib.strerror(0) == "success"
    for code in range(-7, 0):
        assert gswm._lib.strerror(code).startswith("gswm:")
    assert "4-byte" in gswm._lib.strerror(-7) and "16-byte" in gswm._lib.strerror(-7)    # both alignment rules are named
    assert not hasattr(lib, "gswm_workspace_bytes")                    # ABI v1's reserved scratch-memory plumbing is gone
    # argument validation happens before any CUDA call
    assert lib.gswm_embed(None, 0, 0, 0, None, None) == -1
    bad = gswm._lib.Job(1, 1002, 32, 0, 16, 16, 16)
    assert lib.gswm_embed(C.byref(bad), 0, 0, 0, 16, None) == -2
    bad = gswm._lib.Job(1, 16384, 48, 0, 16, 16, 16)
    assert lib.gswm_embed(C.byref(bad), 0, 0, 0, 16, None) == -3        # embed: multiples of 32 only
    assert lib.gswm_embed_mt19937(C.byref(bad), None, 1, 16, 0, None) == -3
    ok = gswm._lib.Job(1, 16384, 256, 0, 16, 16, 16)
    assert lib.gswm_embed(C.byref(ok), 0, 0, -5, 16, None) == -5        # negative global latent index
    assert lib.gswm_embed(C.byref(ok), 0, 1 << 62, 0, 16, None) == -5   # offset >= 2^62
    bad = gswm._lib.Job(1, 16384, 640, 0, 16, 16, 16)           # 640 does not divide 16384
    assert lib.gswm_extract(C.byref(bad), 16, 0, 16, None, None, None, 16, None) == -3
    assert lib.gswm_extract(C.byref(ok), 16, 7, 16, None, None, None, 16, None) == -4
    assert lib.gswm_extract(C.byref(ok), 8, 0, 16, None, None, None, 16, None) == -7
    odd = gswm._lib.Job(1, 36, 12, 0, 16, 16, None)             # 36 fp16 elements = 72-byte rows: not fetchable in 16-byte pieces
    assert lib.gswm_extract(C.byref(odd), 16, 1, 16, None, None, None, 16, None) == -7
    big = gswm._lib.Job(1, 1 << 20, 1 << 14, 0, 16, 16, None)   # messages longer than 8192 bits
    assert lib.gswm_extract(C.byref(big), 16, 0, 16, None, None, None, 16, None) == -5
    assert lib.gswm_chacha20_keystream(16, 16, 1, 100, 16, None) == -2
    assert lib.gswm_debug_philox4x32(16, 4, 8, 16, None) == -5        # 7 or 10 rounds
    # communicator entry points
    h = C.c_void_p()
    assert lib.gswm_comm_create(C.byref(h), 0, 3, 2, None) == gswm._lib.E_COMM       # rank outside [0, n_ranks)
    assert lib.gswm_comm_create(C.byref(h), 0, 0, 33, None) == gswm._lib.E_COMM      # more than GSWM_COMM_MAX_RANKS
    assert lib.gswm_comm_allreduce_counters(None, 16, 4, None) == -1
    assert lib.gswm_allreduce_counters(None, 16, 4, None) == -1


def test_header_is_plain_c_and_links_from_c(gswm, tmp_path):
    """include/gswm.h compiles as strict C99, libgswm.so links from a C program, and argument errors come back before any
    CUDA call (tests/c_abi/abi_check.c runs without a GPU)."""
    exe = tmp_path / "abi_check"
    libdir = os.path.dirname(gswm._lib.LIB_PATH)
    cc = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", f"-I{os.path.join(ROOT, 'include')}",
                         os.path.join(ROOT, "tests", "c_abi", "abi_check.c"), "-o", str(exe), f"-L{libdir}", "-lgswm",
                         f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "abi_check ok" in run.stdout


def test_uniform_source_rounds_match_oracle(gswm):
    # the oracle restates the library's uniform source; both must run the same number of Philox rounds
    assert gswm._lib.lib().gswm_philox_rounds() == O.GSWM_PHILOX_ROUNDS


def test_no_cpu_path(gswm):
    import torch
    km = gswm.KeyMaterial.make(bytes(32), bytes(16), bytes(32), 256)
    with pytest.raises(ValueError):
        gswm.embed_batch(1, (4, 64, 64), km, 0, device="cpu")
    with pytest.raises(ValueError):
        gswm.extract_batch(torch.zeros(1, 4, 64, 64), km)


def test_pad_message_matches_oracle(gswm):
    for msg, nb in [("lthero", 32), ("x" * 40, 32), ("水印-λ-✓", 32), ("ab", 4), (b"\x01\x02", 8), ("lthero", 128)]:
        assert gswm.pad_message(msg, nb) == O.pad_message(msg, nb)
    assert gswm.pad_message("lthero12", 32, use_repeat=True) == O.frame_message("lthero12", 16384, 256, use_repeat=True)[0]
    assert gswm.pad_message("ab", 32, use_repeat=True) == (b"ab" + bytes(6)) * 4
    r1, r2 = gswm.pad_message("", 32), gswm.pad_message("", 32)
    assert len(r1) == 32 and r1 != r2            # empty message -> os.urandom (gs_insert.py:18-20)


def test_resolve_key_nonce(gswm):
    k, n = gswm.resolve_key_nonce(O.DEFAULT_KEY_HEX, O.DEFAULT_NONCE_HEX)
    assert (k, n) == O.resolve_key_nonce(O.DEFAULT_KEY_HEX, O.DEFAULT_NONCE_HEX)
    k, n = gswm.resolve_key_nonce(O.DEFAULT_KEY_HEX, "")
    assert n.hex() == O.DEFAULT_KEY_HEX[16:48] and (k, n) == O.resolve_key_nonce(O.DEFAULT_KEY_HEX, "")
    k, n = gswm.resolve_key_nonce("", "")
    assert len(k) == 32 and len(n) == 16
    with pytest.raises(ValueError):
        gswm.resolve_key_nonce("zz" * 32, "")
    with pytest.raises(ValueError):
        gswm.resolve_key_nonce("00" * 16, "00" * 16)
    with pytest.raises(ValueError):
        gswm.resolve_key_nonce("00" * 32, "00" * 8)


def test_choose_watermark_length(gswm):
    for n in [256, 2047, 2048, 4096, 8192, 16384, 32768, 65536, 1 << 22]:
        assert gswm.choose_watermark_length(n) == O.choose_watermark_length(n)


def test_key_material_shapes(gswm):
    km = gswm.KeyMaterial.make(bytes(32), bytes(16), bytes(32), 256)
    assert not km.per_latent and km.rows == 1
    km = gswm.KeyMaterial.make(bytes(32 * 5), bytes(16 * 5), bytes(32 * 5), 256)
    assert km.per_latent and km.rows == 5
    km = gswm.KeyMaterial.make(bytes(32), bytes(16), bytes(32 * 5), 256)      # shared key, per-latent message
    assert km.rows == 5 and km.keys.shape == (5, 32)
    with pytest.raises(ValueError):
        gswm.KeyMaterial.make(bytes(31), bytes(16), bytes(32), 256)
    with pytest.raises(ValueError):
        gswm.KeyMaterial.make(bytes(32 * 2), bytes(16 * 3), bytes(32), 256)
    with pytest.raises(ValueError):
        gswm.KeyMaterial.make(bytes(32), bytes(16), bytes(32), 100)


def test_calculate_bit_accuracy_matches_reference_vectors(gswm, golden):
    from gswm.extract import calculate_bit_accuracy
    for c in golden["bit_accuracy"]:
        o, a = calculate_bit_accuracy(c["original_message_hex"], c["extracted"])
        assert o == c["original_bin"] and a == c["accuracy"]


def test_two_rank_sharding_and_counter_allreduce_gloo(tmp_path):
    """world_size 2 over gloo on CPU: contiguous batch shards keyed by global latent index reproduce the single-rank
    uniforms, and the all-reduced counters equal the whole-batch counters (the oracle stands in for the kernels)."""
    script = tmp_path / "w.py"
    script.write_text(f"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {ROOT!r}); sys.path.insert(0, os.path.join({ROOT!r}, 'a-watermark-for-diffusion-models_b200'))
from oracle import gs_oracle as O
from gswm.sharding import shard_range
dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
B, n, L, seed = 6, 512, 32, 9
key, nonce = bytes.fromhex(O.DEFAULT_KEY_HEX), bytes.fromhex(O.DEFAULT_NONCE_HEX)
lo, hi = shard_range(B, rank, world)
counters = torch.zeros(4, dtype=torch.int64)
zs = []
for g in range(lo, hi):
    z = O.embed_gswm(b'abcd', key, nonce, seed, 0, g, n, L)
    zs.append(z)
    zn = z + 2.5 * np.random.RandomState(g).standard_normal(n)
    bits = O.recover_message_bits(np.clip(zn, None, 8.0), key, nonce, L)
    m = int((bits == np.unpackbits(np.frombuffer(b'abcd', np.uint8))).sum())
    counters += torch.tensor([m, L, int(m == L), 1])
dist.all_reduce(counters)
gathered = [None] * world
dist.all_gather_object(gathered, np.stack(zs))
if rank == 0:
    np.save({str(tmp_path)!r} + '/z.npy', np.concatenate(gathered))
    np.save({str(tmp_path)!r} + '/c.npy', counters.numpy())
dist.destroy_process_group()
""")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29531", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    z = np.load(tmp_path / "z.npy")
    c = np.load(tmp_path / "c.npy")
    key, nonce = bytes.fromhex(O.DEFAULT_KEY_HEX), bytes.fromhex(O.DEFAULT_NONCE_HEX)
    B, n, L = 6, 512, 32
    whole = np.stack([O.embed_gswm(b"abcd", key, nonce, 9, 0, g, n, L) for g in range(B)])
    assert np.array_equal(z, whole)
    tot = np.zeros(4, dtype=np.int64)
    for g in range(B):
        zn = whole[g] + 2.5 * np.random.RandomState(g).standard_normal(n)
        bits = O.recover_message_bits(np.clip(zn, None, 8.0), key, nonce, L)
        m = int((bits == np.unpackbits(np.frombuffer(b"abcd", np.uint8))).sum())
        tot += [m, L, int(m == L), 1]
    assert np.array_equal(c, tot)


def test_shard_range(gswm):
    from gswm.sharding import shard_range
    for b in [0, 1, 7, 8, 4096, 65537]:
        for w in [1, 2, 3, 8]:
            rs = [shard_range(b, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == b
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in rs]
            assert max(sizes) - min(sizes) <= 1


def test_comfy_nodes_interface_without_comfyui(gswm, monkeypatch):
    """Widget declarations and sampler glue of the ComfyUI drop-in (nodes.py:141-252), against stub comfy modules:
    names, order, defaults and ranges are what the reference declares; the GS latent replaces the start noise."""
    import sys
    import types

    import torch

    from gswm import comfy_nodes as cn

    lat = cn.GSLatent.INPUT_TYPES()["required"]
    assert list(lat) == ["use_seed", "seed", "width", "height", "key", "nonce", "message", "message_length", "batch_size"]
    assert lat["seed"] == ("INT", {"default": 42, "min": 0, "max": 0xffffffff})
    assert lat["width"] == lat["height"] == ("INT", {"default": 512, "min": 64, "max": 8192, "step": 8})
    assert lat["width"][1] is not lat["height"][1]
    assert lat["key"] == ("STRING", {"default": gswm.DEFAULT_KEY_HEX}) and lat["message"] == ("STRING", {"default": "lthero"})
    assert lat["message_length"] == ("INT", {"default": -1, "min": 32, "max": 1024, "step": 32})
    assert lat["batch_size"] == ("INT", {"default": 1, "min": 1, "max": 64})
    ks = cn.GSKSamplerAdvanced.INPUT_TYPES()["required"]
    assert list(ks) == ["model", "add_GS_noise", "add_noise", "noise_seed", "steps", "cfg", "sampler_name", "scheduler",
                        "positive", "negative", "latent_image", "GS_latent_noise", "start_at_step", "end_at_step",
                        "return_with_leftover_noise"]
    assert ks["add_GS_noise"] == (["enable", "disable"],) and ks["add_noise"] == (["disable", "enable"],)
    assert ks["noise_seed"] == ("INT", {"default": 42, "min": 0, "max": 0xffffffffffffffff})
    assert ks["cfg"] == ("FLOAT", {"default": 8.0, "min": 0.0, "max": 100.0, "step": 0.1, "round": 0.01})
    assert ks["GS_latent_noise"] == ("LATENT",) and ks["end_at_step"] == ("INT", {"default": 10000, "min": 0, "max": 10000})
    assert cn.GSLatent.RETURN_TYPES == ("LATENT", "IMAGE") and cn.GSLatent.FUNCTION == "create_gs_latents"
    assert cn.GSKSamplerAdvanced.RETURN_TYPES == ("LATENT",) and cn.GSKSamplerAdvanced.FUNCTION == "sample"

    calls = {}

    def fake_sample(model, noise, steps, cfg, sampler_name, scheduler, positive, negative, latent_image, **kw):
        calls["noise"], calls["kw"], calls["latent_image"] = noise, kw, latent_image
        return noise + 1

    comfy = types.ModuleType("comfy")
    comfy.sample = types.ModuleType("comfy.sample")
    comfy.sample.sample = fake_sample
    comfy.sample.prepare_noise = lambda latent_image, seed, inds: torch.full_like(latent_image, float(seed))
    comfy.utils = types.ModuleType("comfy.utils")
    comfy.utils.PROGRESS_BAR_ENABLED = True
    lp = types.ModuleType("latent_preview")
    lp.prepare_callback = lambda model, steps: ("cb", steps)
    for name, mod in (("comfy", comfy), ("comfy.sample", comfy.sample), ("comfy.utils", comfy.utils), ("latent_preview", lp)):
        monkeypatch.setitem(sys.modules, name, mod)

    empty = {"samples": torch.zeros(2, 4, 8, 8), "noise_mask": "mask"}
    gs = {"samples": torch.ones(2, 4, 8, 8)}
    node = cn.GSKSamplerAdvanced()
    (out,) = node.sample("m", "enable", "enable", 7, 20, 8.0, "euler", "normal", "p", "n", empty, gs, 0, 10000, "disable")
    assert torch.equal(calls["noise"], gs["samples"]) and torch.equal(out["samples"], gs["samples"] + 1)
    assert out["noise_mask"] == "mask" and out is not empty and torch.equal(empty["samples"], torch.zeros(2, 4, 8, 8))
    kw = calls["kw"]
    assert kw["force_full_denoise"] is True and kw["disable_noise"] is False and kw["noise_mask"] == "mask"
    assert kw["callback"] == ("cb", 20) and kw["disable_pbar"] is False and kw["seed"] == 7
    assert kw["start_step"] == 0 and kw["last_step"] == 10000 and kw["denoise"] == 1.0
    node.sample("m", "disable", "enable", 7, 20, 8.0, "euler", "normal", "p", "n", empty, gs, 0, 10000, "enable")
    assert torch.equal(calls["noise"], torch.full((2, 4, 8, 8), 7.0)) and calls["kw"]["force_full_denoise"] is False
    node.sample("m", "disable", "disable", 7, 20, 8.0, "euler", "normal", "p", "n", empty, gs, 0, 10000, "enable")
    assert torch.equal(calls["noise"], torch.zeros(2, 4, 8, 8)) and calls["noise"].device.type == "cpu"
    assert calls["kw"]["disable_noise"] is True


def test_comfy_unseeded_batch_host_logic_follows_the_reference(gswm, golden, monkeypatch, tmp_path):
    """Host side of GSLatent(use_seed=0) without a GPU: the device call is replaced by a recorder.  Per latent -- as the
    reference's batch_size sequential calls do (nodes.py:236-237) -- a fresh os.urandom message / key / nonce when the
    widgets are empty, its own info_data.txt record, and the widget's seed in the log (nodes.py:131,134)."""
    import importlib.util
    import os

    import torch

    from gswm import _embed_common as common
    from gswm import comfy_nodes as cn

    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    seen = {}

    def fake_embed(u, latent_shape, key, nonce, k, msg_bits, n_latents, out_dtype=np.float64):
        seen.update(u=u, shape=latent_shape, key=key, nonce=nonce, k=k, bits=msg_bits, n=n_latents)
        return np.zeros((n_latents, *latent_shape), dtype=out_dtype)

    monkeypatch.setattr(common, "embed_injected_host", fake_embed)
    monkeypatch.chdir(tmp_path)
    g = golden["gslatent_unseeded_random"]
    monkeypatch.setattr(os, "urandom", mg._FakeUrandom())
    np.random.seed(g["np_seed"])
    lat, first = cn.GSLatent().create_gs_latents("", "", "", g["batch_size"], 0, g["widget_seed"], 256, 256, -1)
    assert list(lat["samples"].shape) == g["shape"] and seen["bits"] == 128 and seen["n"] == 2
    assert len(seen["key"]) == 64 and len(seen["nonce"]) == 32 and len(seen["k"]) == 32      # one row per latent
    assert np.array_equal(seen["u"], np.random.RandomState(g["np_seed"]).uniform(size=(2, 4096)))
    lines = (tmp_path / "info_data.txt").read_text().splitlines()
    want = g["info_data_tail"]
    assert [a for a in lines if not a.startswith("Time: ")] == [b for b in want if not b.startswith("Time: ")]
    # nothing random: one shared row, still one record per latent
    (tmp_path / "info_data.txt").unlink()
    g = golden["gslatent_unseeded"]
    cn.GSLatent().create_gs_latents(gswm.DEFAULT_KEY_HEX, gswm.DEFAULT_NONCE_HEX, "lthero", 3, 0, 77, 512, 512, 256)
    assert len(seen["key"]) == 32 and len(seen["k"]) == 32 and seen["n"] == 3
    lines = (tmp_path / "info_data.txt").read_text().splitlines()
    assert [a for a in lines if not a.startswith("Time: ")] == [b for b in g["info_data_tail"] if not b.startswith("Time: ")]


def test_webui_scripts_patch_and_restore_with_stub_webui(gswm, monkeypatch):
    """Script.run of both webui drop-ins against stub ``modules`` / ``gradio``: the hook is in place while
    process_images runs (v1.5.2:123-138, v1.6.0:175-190), the UI fields land in the shared globals, and the original
    hook is back afterwards -- also when process_images raises (the reference's v1.6.0 never restores it)."""
    import sys
    import types

    from gswm import webui_v152 as w5
    from gswm import webui_v160 as w6

    seen = {}
    processing = types.ModuleType("modules.processing")
    processing.create_random_tensors = original_creator = lambda *a, **k: "stock tensors"
    rng = types.ModuleType("modules.rng")
    rng.ImageRNG = original_rng = type("ImageRNG", (), {})

    def process_images(p):
        seen["creator"], seen["rng"] = processing.create_random_tensors, rng.ImageRNG
        seen["globals"] = (w5.global_message, w5.global_key, w5.global_nonce, w5.global_randomSeed,
                           w5.global_use_randomSeed, w5.global_use_repeat)
        if p == "boom":
            raise RuntimeError("sampler failed")
        return "processed"

    processing.process_images = process_images
    scripts = types.ModuleType("modules.scripts")
    scripts.Script = type("Script", (), {})
    modules = types.ModuleType("modules")
    modules.processing, modules.scripts, modules.rng = processing, scripts, rng
    gradio = types.ModuleType("gradio")
    for name, mod in (("modules", modules), ("modules.processing", processing), ("modules.scripts", scripts),
                      ("modules.rng", rng), ("gradio", gradio)):
        monkeypatch.setitem(sys.modules, name, mod)

    s5 = w5._make_script()()
    assert s5.title() == "GS_watermark_insert"
    assert s5.run("p", "msg", "aa" * 32, "bb" * 16, "1234", "1", "0") == "processed"
    assert seen["creator"] is w5.advanced_creator and processing.create_random_tensors is original_creator
    assert seen["globals"] == ("msg", "aa" * 32, "bb" * 16, 1234, 1, 0)
    with pytest.raises(RuntimeError):
        s5.run("boom", "m", "k", "n", "5", "0", "1")
    assert processing.create_random_tensors is original_creator and seen["globals"][3:] == (5, 0, 1)

    s6 = w6._make_script()()
    assert s6.run("p", "other", "cc" * 32, "", "77", "0", "1") == "processed"
    assert seen["rng"] is w6.modified_ImageRNG and rng.ImageRNG is original_rng
    assert seen["globals"] == ("other", "cc" * 32, "", 77, 0, 1)
    assert (w6.global_message, w6.global_randomSeed) == ("other", 77)        # one shared state for both generations
    with pytest.raises(RuntimeError):
        s6.run("boom", "m", "k", "n", "5", "0", "0")
    assert rng.ImageRNG is original_rng


def test_bench_clock_sampler_selects_samples_inside_the_timed_window():
    """bench.py's nvidia-smi sampler: lines are parsed by timestamp, only those inside the timed window count (the two
    nearest ones if none falls inside), throttle reasons are collected from the selected lines."""
    import datetime as dt

    sys.path.insert(0, ROOT)
    import bench

    class Done:
        def terminate(self): pass
        def wait(self, timeout=None): return 0
        def kill(self): pass

    def sampler_with(lines):
        s = bench.ClockSampler(0)
        with open(s.path, "w") as f:
            f.write("\n".join(lines) + "\n")
        s.proc = Done()
        return s

    lines = ["2026/10/17 12:00:00.000, 1965, 1965, 200.0, 0x0000000000000000, Not Active, Not Active, Not Active, Not Active",
             "2026/10/17 12:00:00.020, 1900, 1965, 900.5, 0x0000000000000004, Not Active, Not Active, Not Active, Active",
             "2026/10/17 12:00:00.040, 1700, 1965, 1001.0, 0x0000000000000004, Not Active, Not Active, Not Active, Active",
             "garbage line",
             "2026/10/17 12:00:00.060, 1965, 1965, 150.0, 0x0000000000000000, Not Active, Active, Not Active, Not Active"]
    t = lambda ms: dt.datetime(2026, 10, 17, 12, 0, 0, ms * 1000)
    got = sampler_with(lines).stop((t(10), t(50)))
    assert got["samples"] == 2 and got["samples_total"] == 4 and got["sm_mhz"] == 1800.0 and got["sm_max_mhz"] == 1965.0
    assert got["reasons"] == ["sw_power_cap"] and got["power_w_max"] == 1001.0 and got["window"].startswith("timed region")
    got = sampler_with(lines).stop((t(100), t(200)))                     # nothing inside: the two nearest samples, and says so
    assert got["samples"] == 2 and got["window"].startswith("the two samples nearest") and got["sm_mhz"] == 1832.5
    assert got["reasons"] == ["hw_thermal_slowdown", "sw_power_cap"]
    got = sampler_with(lines).stop((t(21), t(23)))                       # a window between two samples: its neighbours
    assert got["samples"] == 2 and got["sm_mhz"] == 1800.0 and got["reasons"] == ["sw_power_cap"]
    got = sampler_with(lines).stop(None)                                 # no window at all: every sample
    assert got["samples"] == 4 and got["window"] == "whole run"
    assert sampler_with(["garbage"]).stop(None)["reasons"] == ["no samples"]
    s = bench.ClockSampler(0)
    assert s.stop()["reasons"] == ["nvidia-smi unavailable"] and s.wait_ready(0.01) is False


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours): one JSON line with the contract's keys."""
    import json

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--ref-step-seconds", "0.2"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "latents/s" and d["higher_is_better"] is True and d["steps"] == 2
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["config"]["decoded_ok"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "latents/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["gpu_launches"] == 0
    # under torchrun only rank 0 works; the others exit 0 without output
    other = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1"],
                           capture_output=True, text=True, timeout=60, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert other.returncode == 0 and other.stdout.strip() == ""


def test_bench_gpu_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback anywhere: without a CUDA device the product arm of bench.py exits non-zero and prints no JSON."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
    assert not any(ln.lstrip().startswith("{") for ln in out.stdout.splitlines())


def test_kernel_coefficients_are_what_the_fit_tool_derives(tmp_path):
    """csrc/gswm_coeffs.inc is generated, not transcribed: tools/fit_halfnormal_quantile.py with the parameters recorded in
    the file's header reproduces it byte for byte, and reports an emulated-fp32 error inside the 1e-6 tolerance."""
    inc = os.path.join(ROOT, "a-watermark-for-diffusion-models_b200", "csrc", "gswm_coeffs.inc")
    header = open(inc).read().splitlines()[1]
    m = re.match(r"// deg_central=(\d+) deg_tail=(\d+) w_split=([\d.]+)", header)
    assert m, header
    deg64 = len(re.search(r"#define GSWM_HNQ64_CENTRAL_COEFFS (.*)", open(inc).read()).group(1).split(",")) - 1
    out = tmp_path / "coeffs.inc"
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fit_halfnormal_quantile.py"), "--deg-central", m.group(1),
                          "--deg-tail", m.group(2), "--w-split", m.group(3), "--deg64", str(deg64), "--out", str(out)],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    assert out.read_text() == open(inc).read()
    errs = [float(x) for x in re.search(r"max rel err central ([\d.e+-]+)\s+tail ([\d.e+-]+)", res.stdout).groups()]
    assert max(errs) < 6e-7
