"""Pin the oracle (oracle/gs_oracle.py) against vectors produced by the reference itself.

The fixtures in tests/golden/ were written by tests/golden/make_golden.py, which imports and runs
the unmodified reference (gs_insert.py, nodes.py, the webui script, extract.py) in the build
container.  CPU only.
"""
import hashlib

import numpy as np
import pytest

from oracle import gs_oracle as O

KEY = bytes.fromhex(O.DEFAULT_KEY_HEX)
NONCE = bytes.fromhex(O.DEFAULT_NONCE_HEX)


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def signs(z) -> np.ndarray:
    return np.packbits((np.asarray(z, dtype=np.float64).reshape(-1) >= 0).astype(np.uint8))


# ---------------------------------------------------------------- ChaCha20
def test_chacha20_matches_reference_library(golden):
    for c in golden["chacha20"]:
        s = O.chacha20_keystream(bytes.fromhex(c["key"]), bytes.fromhex(c["nonce"]), c["nbytes"]).tobytes()
        assert s[:64].hex() == c["first64"], c["name"]
        assert s[-64:].hex() == c["last64"], c["name"]
        assert hashlib.sha256(s).hexdigest() == c["sha256"], c["name"]


def test_chacha20_rfc7539_block():
    # RFC 7539 section 2.3.2 keystream block (counter 1, nonce 00 00 00 09 00 00 00 4a 00 00 00 00)
    s = O.chacha20_keystream(bytes(range(32)), bytes.fromhex("01000000000000090000004a00000000"), 64)
    assert s.tobytes().hex().startswith("10f1e7e4d13b5915500fdd1fa32071c4c7d1f4c733c068030422aa9ac3d46c4e")


def test_chacha20_live_library_agrees():
    rs = np.random.RandomState(11)
    for _ in range(5):
        k, n = rs.bytes(32), rs.bytes(16)
        nb = int(rs.randint(1, 5000))
        assert np.array_equal(O.chacha20_keystream(k, n, nb), O.chacha20_keystream_lib(k, n, nb))


def test_chacha20_counter_is_64_bit():
    a = O.chacha20_blocks(KEY, bytes.fromhex("ffffffff000000000102030405060708"), 2)
    b = O.chacha20_blocks(KEY, bytes.fromhex("00000000010000000102030405060708"), 1)
    assert np.array_equal(a[1], b[0])  # carry into word 13


# ---------------------------------------------------------------- embed
def test_embed_cli_matches_reference(golden, golden_arrays):
    for c in golden["embed_cli"]:
        key, nonce = O.resolve_key_nonce(c["key_hex"], c["nonce_hex"])
        u = np.random.RandomState(c["u_seed"]).uniform(size=16384)
        z = O.embed(c["message"], key, nonce, u, 256).reshape(4, 64, 64)
        assert sha(z) == c["sha256_f64"], c["name"]
        assert sha(z.astype(np.float32)) == c["sha256_f32"], c["name"]
        assert sha(signs(z)) == c["sha256_signs"], c["name"]
        assert np.array_equal(z.reshape(-1)[:512], golden_arrays[c["name"] + "_z64_head"])
    assert np.array_equal(
        O.embed("lthero", KEY, NONCE, np.random.RandomState(1234).uniform(size=16384)).astype(np.float32),
        golden_arrays["cli_lthero_z32"].reshape(-1))


def test_embed_comfy_matches_reference(golden, golden_arrays):
    for c in golden["embed_comfy"]:
        n = 4 * (c["width"] // 8) * (c["height"] // 8)
        L = c["message_length"] if c["message_length"] != -1 else O.choose_watermark_length(n)
        u = np.random.RandomState(c["seed"]).uniform(size=n)
        z = O.embed(c["message"], KEY, NONCE, u, L).astype(np.float32)
        assert sha(z) == c["sha256_f32"], c["name"]
        assert sha(signs(z)) == c["sha256_signs"], c["name"]
        assert np.array_equal(z[:256], golden_arrays[c["name"] + "_z32_head"])


def _fake_urandom():
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location(
        "make_golden", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod._FakeUrandom()


def test_embed_comfy_unseeded_batch_matches_reference(golden, golden_arrays):
    """GSLatent with use_seed=0 (nodes.py:236-237): batch_size sequential calls drawing from numpy's global stream; with an
    empty message / key each call draws its own os.urandom message, key and nonce (nodes.py:76,97-98)."""
    g = golden["gslatent_unseeded"]
    u = np.random.RandomState(g["np_seed"]).uniform(size=(g["batch_size"], 16384))     # np.random.seed(s) == RandomState(s) stream
    z = np.stack([O.embed("lthero", KEY, NONCE, u[i], 256).astype(np.float32) for i in range(g["batch_size"])])
    assert sha(z.reshape(g["shape"])) == g["sha256"]
    assert [sha(signs(z[i])) for i in range(g["batch_size"])] == g["sha256_signs_each"]
    assert np.array_equal(z[:, :128], golden_arrays["gslatent_unseeded_heads"])

    g = golden["gslatent_unseeded_random"]
    fake = _fake_urandom()
    n = 4 * 32 * 32
    L = O.choose_watermark_length(n)
    u = np.random.RandomState(g["np_seed"]).uniform(size=(g["batch_size"], n))
    zs, recs = [], []
    for i in range(g["batch_size"]):
        k, key, nonce = fake(L // 8), fake(32), fake(16)                              # the reference's call order
        zs.append(O.embed(k, key, nonce, u[i], L).astype(np.float32))
        recs.append((key.hex(), nonce.hex(), k.hex()))
    z = np.stack(zs)
    assert sha(z.reshape(g["shape"])) == g["sha256"]
    assert np.array_equal(z[:, :128], golden_arrays["gslatent_unseeded_random_heads"])
    tail = g["info_data_tail"]
    for i, (kh, nh, mh) in enumerate(recs):
        assert tail[10 * i + 1:10 * i + 4] == [f"key: {kh}", f"nonce: {nh}", f"message: {mh}"]
        assert tail[10 * i + 4] == f"randomSeed: {g['widget_seed']}"


def test_embed_webui_matches_reference(golden):
    for c in golden["embed_webui"]:
        u = np.random.RandomState(c["seed"]).uniform(size=16384)
        z = O.embed(c["message"], KEY, NONCE, u, 256, use_repeat=bool(c["use_repeat"]))
        assert sha(z) == c["sha256_f64"], c["name"]


def test_embed_scalar_form_equals_vectorised():
    u = np.random.RandomState(3).uniform(size=1024)
    a = np.array(O.embed_scalar("ab", KEY, NONCE, u, 32))
    b = O.embed("ab", KEY, NONCE, u, 32)
    assert np.array_equal(a, b)


def test_ppf_edges(golden):
    for e in golden["ppf_edges"]:
        p = float(e["p"])
        z = float(O.ndtri(p))
        assert repr(z) == e["z"]


def test_survey_known_answers():
    # encrypted tile head and keystream head for the default key / nonce
    assert O.chacha20_keystream(KEY, NONCE, 16).tobytes().hex() == "610848b70a836027e692a131b12cdcd9"
    _, s_d = O.frame_message("lthero", 16384, 256)
    y = O.bucket_bits(s_d, KEY, NONCE)
    assert np.packbits(y)[:8].tobytes().hex() == "0d7c20d278ec6027"
    assert sha(np.packbits(y)) == "055b3611184c4b7cfdc6ab5cb7d4c0476d76eb2232cc9a48e5788f95a9e13cb7"


# ---------------------------------------------------------------- extract
def _noisy(base, sigma, seed, dtype):
    zn = base.astype(np.float64)
    if sigma:
        zn = zn + sigma * np.random.RandomState(seed).standard_normal(zn.shape)
    return np.clip(zn, -60000.0, 8.0).astype(dtype)


def test_extract_matches_reference(golden, golden_arrays):
    base = golden_arrays["cli_lthero_z32"]
    for c in golden["extract"]:
        if "noise_seed" not in c:
            continue
        z = _noisy(base, c["sigma"], c["noise_seed"], c["dtype"])
        got = O.recover_message(z, KEY, NONCE, 256)
        assert got == c["extracted_bin"], c["name"]
        orig, acc = O.calculate_bit_accuracy((b"lthero" + bytes(26)).hex(), got)
        assert acc == c["bit_accuracy"] and orig == c["original_bin"]


def test_extract_other_lengths_match_reference(golden):
    for c in golden["extract"]:
        if "width" not in c:
            continue
        n = 4 * (c["width"] // 8) * (c["height"] // 8)
        L = c["message_length"]
        u = np.random.RandomState(c["seed"]).uniform(size=n)
        z = O.embed(c["message"], KEY, NONCE, u, L).astype(np.float32)
        zn = (z.astype(np.float64).reshape(4, c["height"] // 8, c["width"] // 8)
              + 1.5 * np.random.RandomState(c["seed"]).standard_normal((4, c["height"] // 8, c["width"] // 8)))
        got = O.recover_message(zn.astype(np.float16), KEY, NONCE, L)
        assert got == c["extracted_bin"], c["name"]


def test_extract_scalar_form_equals_vectorised():
    z = np.random.RandomState(5).standard_normal(2048).astype(np.float32)
    assert O.recover_message_scalar(z, KEY, NONCE, 64) == O.recover_message(z, KEY, NONCE, 64)


def test_quantise_edges(golden):
    for e in golden["quantise_edges"]:
        z = np.array([float(e["z"])] * 8, dtype=np.float64)
        assert int(O.quantise(z)[0]) == e["bit"], e
    # the closed form the CUDA kernel uses
    for e in golden["quantise_edges"]:
        assert int(float(e["z"]) >= O.CDF_HALF_THRESHOLD) == e["bit"], e


def test_quantise_raises_like_reference(golden):
    for e in golden["quantise_raises"]:
        assert e["raises"] == "ValueError"
        with pytest.raises(ValueError):
            O.quantise(np.array([float(e["z"])] + [0.5] * 7))
    assert int(O.quantise(np.array([np.nextafter(O.CDF_ONE_THRESHOLD, 0)] * 8))[0]) == 1


def test_bit_accuracy_cases(golden):
    for c in golden["bit_accuracy"]:
        o, a = O.calculate_bit_accuracy(c["original_message_hex"], c["extracted"])
        assert o == c["original_bin"] and a == c["accuracy"]


def test_round_trip_all_lengths():
    for n, L in [(512, 32), (1024, 64), (16384, 256), (16384, 512), (65536, 1024), (24576, 96)]:
        u = np.random.RandomState(n + L).uniform(size=n)
        msg = bytes(np.random.RandomState(L).randint(1, 256, size=L // 8).astype(np.uint8))
        z = O.embed(msg, KEY, NONCE, u, L)
        assert O.bits_to_bytes(O.recover_message_bits(z, KEY, NONCE, L)) == msg


# ---------------------------------------------------------------- the product's uniform source
def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    z = O.philox4x32(np.zeros((1, 4), np.uint32), (0, 0))[0]
    assert [hex(int(v)) for v in z] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    f = O.philox4x32(np.full((1, 4), 0xFFFFFFFF, np.uint32), (0xFFFFFFFF, 0xFFFFFFFF))[0]
    assert [hex(int(v)) for v in f] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    p = O.philox4x32(np.array([[0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344]], np.uint32),
                     (0xA4093822, 0x299F31D0))[0]
    assert [hex(int(v)) for v in p] == ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_gswm_uniform_source():
    m = O.gswm_uniform_ints(0x5EED, 0, 3, 16384)
    assert m.dtype == np.uint32 and m.size == 16384 and int(m.max()) < (1 << 23)
    # a shorter latent is NOT a prefix (the counter depends on tiles per latent) but is deterministic
    assert np.array_equal(O.gswm_uniform_ints(0x5EED, 0, 3, 16384), m)
    assert not np.array_equal(O.gswm_uniform_ints(0x5EED, 1, 3, 16384), m)
    assert not np.array_equal(O.gswm_uniform_ints(0x5EED, 0, 4, 16384), m)
    # all four float4 of a super-iteration come from different bits
    assert len(np.unique(m[:64])) > 60
    y = np.random.RandomState(0).randint(0, 2, size=16384)
    u = O.gswm_uniforms(0x5EED, 0, 3, y)
    assert u.min() > 0 and u.max() < 1 and abs(u.mean() - 0.5) < 0.01
    v = (m + 0.5) * 2.0 ** -23
    assert np.array_equal(u[y == 1], v[y == 1]) and np.array_equal(u[y == 0], 1 - v[y == 0])
    # uniformity of the 23-bit integers: chi-square over 256 buckets of the top byte
    big = np.concatenate([O.gswm_uniform_ints(7, 0, i, 65536) for i in range(8)])
    hist = np.bincount(big >> 15, minlength=256)
    chi2 = ((hist - big.size / 256) ** 2 / (big.size / 256)).sum()
    assert chi2 < 340          # 255 dof: P(chi2 > 340) ~ 3e-4
    z = O.embed_gswm("lthero", KEY, NONCE, 0x5EED, 0, 0, 16384)
    assert O.bits_to_bytes(O.recover_message_bits(z, KEY, NONCE, 256))[:6] == b"lthero"
    # uniforms v3: the outermost cell is subdivided (latent 404 of seed 0x5EED holds one such element, number 503)
    m = O.gswm_uniform_ints(0x5EED, 0, 404, 16384)
    assert m[503] == O.GSWM_TOP_CELL and (m == O.GSWM_TOP_CELL).sum() == 1
    w = int(O.gswm_top_cell_words(0x5EED, 0, 404, 16384, [503])[0])
    for ybit in (0, 1):
        u = O.gswm_uniforms(0x5EED, 0, 404, np.full(16384, ybit))
        tail = ((w >> 4) + 0.5) * 2.0 ** -51                      # 1 - v: exact in float64
        assert u[503] == (1.0 - tail if ybit else tail) and 0 < tail < 2.0 ** -23
        zz = O.embed_from_uniform(np.full(16384, ybit), u)
        assert 5.29 < abs(zz[503]) <= 8.2096 and (zz[503] > 0) == bool(ybit)
        others = np.delete(np.arange(16384), 503)
        v = (m[others] + 0.5) * 2.0 ** -23
        assert np.array_equal(u[others], v if ybit else 1 - v)


def test_truncated_tail_mass_is_stated():
    """What the uniform grid can and cannot produce (VERDICT weak #7): with the outermost cell refined the support is
    |z| in [7.5e-8, 8.2095]; the mass beyond it is 2^-52 (2.2e-16) per element, of the order of the reference's own 2^-53.
    Without the refinement the outermost cell would collapse to its midpoint, |z| <= 5.42, cutting off a tail of mass
    2^-24 = 6.0e-8 per element (one element in 16.8 M: about four per 4096-latent batch)."""
    from scipy.special import ndtri
    from scipy.stats import norm
    assert abs(float(ndtri(0.5 + 0.5 * (0.5 * 2.0 ** -23))) - 7.4703e-8) < 1e-11          # smallest |z| (m = 0)
    zmax_v2 = float(ndtri(1.0 - 0.5 * 0.5 * 2.0 ** -23))                                  # m = 2^23 - 1 at its midpoint
    assert abs(zmax_v2 - 5.42) < 5e-3 and abs(2 * norm.sf(zmax_v2) - 2.0 ** -24) < 1e-12  # P(|Z| > 5.42) = 2^-24 = 6.0e-8
    zmax_v3 = float(-ndtri(0.5 * 2.0 ** -52))                                             # m2 = 0
    assert abs(zmax_v3 - 8.2095) < 1e-4 and zmax_v3 < O.CDF_ONE_THRESHOLD                 # below what extract.py can parse
    assert abs(2 * norm.sf(zmax_v3) / 2.0 ** -52 - 1) < 1e-9                              # tail mass cut off by v3
