"""Parity of the CUDA path (libgswm.so through its C ABI) against the oracle and the golden vectors.

Run on the B200 box:  python -m pytest tests -m gpu -x -q
Bars (BASELINE.json north_star): keystream, bucket membership (sign), vote counts, decoded messages and
bit counts BIT-EXACT; latents within 1e-6 relative of scipy's float64 norm.ppf.
"""
import ctypes as C
import hashlib
import os
import types

import numpy as np
import pytest
import torch

from oracle import gs_oracle as O

pytestmark = pytest.mark.gpu

REL_TOL = 1e-6   # north_star: latents within 1e-6 relative of scipy's float64 norm.ppf
KEY = bytes.fromhex(O.DEFAULT_KEY_HEX)
NONCE = bytes.fromhex(O.DEFAULT_NONCE_HEX)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gswm(cuda_device):
    import gswm as g
    g.build()      # no-op when libgswm.so is up to date (it travels with the tree); compiles it with nvcc otherwise
    g._lib.lib()   # raises if the extension cannot be loaded: there is no fallback path
    return g


def rel_err(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        e = np.abs(got - ref) / np.abs(ref)
    e[(ref == 0) & (got == 0)] = 0.0
    e[np.isinf(ref) & (got == ref)] = 0.0
    return e


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ------------------------------------------------------------------------------------ K1 ChaCha20
def test_chacha20_golden(gswm, golden):
    for c in golden["chacha20"]:
        ks = gswm.chacha20_keystream(bytes.fromhex(c["key"]), bytes.fromhex(c["nonce"]), c["nbytes"]).cpu().numpy()[0]
        s = ks.tobytes()
        assert s[:64].hex() == c["first64"], c["name"]
        assert s[-64:].hex() == c["last64"], c["name"]
        assert hashlib.sha256(s).hexdigest() == c["sha256"], c["name"]


def test_chacha20_many_keys_vs_oracle(gswm):
    rs = np.random.RandomState(2025)
    n = 300
    keys = np.frombuffer(rs.bytes(32 * n), np.uint8).reshape(n, 32).copy()
    nonces = np.frombuffer(rs.bytes(16 * n), np.uint8).reshape(n, 16).copy()
    nonces[0, :8] = 0xFF                       # 64-bit counter wraps to zero
    nonces[1, :4] = 0xFF; nonces[1, 4:8] = 0    # carry into word 13
    got = gswm.chacha20_keystream(keys, nonces, 2048).cpu().numpy()
    for i in range(n):
        ref = O.chacha20_keystream(keys[i].tobytes(), nonces[i].tobytes(), 2048)
        assert np.array_equal(got[i], ref), i
    lib_ref = O.chacha20_keystream_lib(keys[1].tobytes(), nonces[1].tobytes(), 2048)
    assert np.array_equal(got[1], lib_ref)


# ------------------------------------------------------------------------------------ fp32 quantile, exhaustive
@pytest.mark.parametrize("vec4", [0, 1])
def test_bucket_quantile_exhaustive(gswm, cuda_device, vec4):
    """All 2^23 uniforms x both buckets: sign exact, value within 1e-6 relative of float64 ndtri."""
    lib = gswm._lib.lib()
    worst = 0.0
    for bucket in (0, 1):
        for lo in range(0, 1 << 23, 1 << 21):
            m = np.arange(lo, lo + (1 << 21), dtype=np.uint32)
            words = (m << np.uint32(9)) | np.uint32(0x1A5)            # low 9 bits are ignored by the kernel
            d_w = torch.from_numpy(words.view(np.int32)).to(cuda_device)
            d_o = torch.empty(words.size, dtype=torch.float32, device=cuda_device)
            rc = lib.gswm_debug_bucket_quantile(d_w.data_ptr(), words.size, bucket, vec4, d_o.data_ptr(), None)
            assert rc == 0
            torch.cuda.synchronize()
            got = d_o.cpu().numpy()
            v = (m.astype(np.float64) + 0.5) * 2.0 ** -23
            u = v if bucket else 1.0 - v                  # the kernel's uniform for bucket 0 is the complement
            ref = O.embed_from_uniform(np.full(u.shape, bucket), u)
            assert np.array_equal(got >= 0, ref >= 0), "bucket membership must be bit-exact"
            assert np.array_equal(got >= 0, np.full(u.shape, bool(bucket)))
            worst = max(worst, float(rel_err(got, ref).max()))
    print(f"exhaustive max relative error (vec4={vec4}): {worst:.3e}")
    assert worst <= REL_TOL


def test_norm_ppf_f64(gswm, cuda_device):
    lib = gswm._lib.lib()
    rs = np.random.RandomState(1)
    p = np.concatenate([rs.uniform(size=200000), rs.uniform(size=50000) * 1e-6, 1 - rs.uniform(size=50000) * 1e-9,
                        2.0 ** -np.arange(1, 1070, dtype=np.float64), [0.0, 0.5, 1.0, 1 - 2.0 ** -53, 2.0 ** -54, 0.25],
                        0.5 + (rs.uniform(size=1000) - 0.5) * 1e-12])
    d_p = torch.from_numpy(p).to(cuda_device)
    d_o = torch.empty_like(d_p)
    assert lib.gswm_debug_norm_ppf(d_p.data_ptr(), p.size, d_o.data_ptr(), None) == 0
    got = d_o.cpu().numpy()
    ref = O.ndtri(p)
    assert np.array_equal(np.isinf(got), np.isinf(ref))
    assert np.array_equal(np.signbit(got[ref != 0]), np.signbit(ref[ref != 0]))
    e = rel_err(got, ref)
    print("fp64 ppf max rel err:", e.max())
    assert e.max() <= 1e-9
    # outside [0, 1] (an injected "uniform" that is none): nan, as scipy's norm.ppf answers -- not +-inf
    bad = np.array([-1e-300, -0.25, 1.0000000000000002, 2.5, np.nan, -np.inf, np.inf])
    d_b = torch.from_numpy(bad).to(cuda_device)
    d_bo = torch.empty_like(d_b)
    assert lib.gswm_debug_norm_ppf(d_b.data_ptr(), bad.size, d_bo.data_ptr(), None) == 0
    assert np.isnan(d_bo.cpu().numpy()).all() and np.isnan(O.ndtri(bad)).all()


# ------------------------------------------------------------------------------------ K2 embed
def oracle_embed_batch(messages, keys, nonces, n, L, seed, offset, first_latent, b):
    out = np.empty((b, n), dtype=np.float64)
    for i in range(b):
        k = keys[i] if isinstance(keys, list) else keys
        no = nonces[i] if isinstance(nonces, list) else nonces
        m = messages[i] if isinstance(messages, list) else messages
        out[i] = O.embed_gswm(m, k, no, seed, offset, first_latent + i, n, L)
    return out


@pytest.mark.parametrize("shape,L", [((4, 64, 64), 256), ((4, 128, 128), 256), ((4, 96, 64), 96), ((4, 8, 16), 32),
                                     ((4, 64, 64), 1024), ((4, 72, 64), 512), ((4, 160, 128), 320), ((4, 8, 8), 32),
                                     ((4, 152, 104), 256), ((4, 9, 9), 32), ((4, 72, 72), 1024)])
def test_embed_shared_key_vs_oracle(gswm, cuda_device, shape, L):
    n = int(np.prod(shape))
    b, seed, offset, first = 5, 0x5EED, (7 << 32) + 3, 1000
    msg = bytes(np.random.RandomState(L).randint(0, 256, size=L // 8).astype(np.uint8))
    km = gswm.KeyMaterial.make(KEY, NONCE, msg, L)
    z = gswm.embed_batch(b, shape, km, seed, offset, first, cuda_device).cpu().numpy().reshape(b, n)
    ref = oracle_embed_batch(msg, KEY, NONCE, n, L, seed, offset, first, b)
    assert np.array_equal(z >= 0, ref >= 0), "bucket membership must be bit-exact"
    assert rel_err(z, ref).max() <= REL_TOL
    assert np.isfinite(z).all()


def test_embed_per_latent_keys_vs_oracle(gswm, cuda_device):
    rs = np.random.RandomState(77)
    b, shape, L = 37, (4, 64, 64), 256
    n = 16384
    keys = [rs.bytes(32) for _ in range(b)]
    nonces = [rs.bytes(16) for _ in range(b)]
    msgs = [rs.bytes(32) for _ in range(b)]
    km = gswm.KeyMaterial.make(b"".join(keys), b"".join(nonces), b"".join(msgs), L)
    assert km.per_latent
    z = gswm.embed_batch(b, shape, km, 99, 0, 0, cuda_device).cpu().numpy().reshape(b, n)
    ref = oracle_embed_batch(msgs, keys, nonces, n, L, 99, 0, 0, b)
    assert np.array_equal(z >= 0, ref >= 0)
    assert rel_err(z, ref).max() <= REL_TOL


def test_embed_sharding_is_transparent(gswm, cuda_device):
    km = gswm.KeyMaterial.make(KEY, NONCE, gswm.pad_message("lthero", 32), 256)
    whole = gswm.embed_batch(16, (4, 64, 64), km, 7, 0, 0, cuda_device)
    parts = [gswm.embed_batch(4, (4, 64, 64), km, 7, 0, 4 * r, cuda_device) for r in range(4)]
    assert torch.equal(whole, torch.cat(parts))


def test_embed_injected_golden(gswm, cuda_device, golden, golden_arrays):
    """Same injected uniforms as the reference run that produced the golden latent."""
    u = np.random.RandomState(1234).uniform(size=16384)
    km = gswm.KeyMaterial.make(KEY, NONCE, gswm.pad_message("lthero", 32), 256)
    z64 = gswm.embed_batch_injected(torch.from_numpy(u).to(cuda_device), (4, 64, 64), km, 1, torch.float64).cpu().numpy()
    z32 = gswm.embed_batch_injected(torch.from_numpy(u).to(cuda_device), (4, 64, 64), km, 1, torch.float32).cpu().numpy()
    gold32 = golden_arrays["cli_lthero_z32"]
    ref64 = O.embed("lthero", KEY, NONCE, u, 256).reshape(1, 4, 64, 64)
    assert np.array_equal(z64 >= 0, ref64 >= 0)
    assert rel_err(z64, ref64).max() <= 1e-9
    assert rel_err(z32, gold32[None]).max() <= REL_TOL
    c = [e for e in golden["embed_cli"] if e["name"] == "cli_lthero"][0]
    packed = np.packbits((z32.reshape(-1) >= 0).astype(np.uint8))
    assert sha(packed) == c["sha256_signs"]


def test_config2_injected_parity_subset_64_latents(gswm, cuda_device):
    """the first 64 samples re-run in injected-uniform mode with u = RandomState(1234 + b).uniform
    (float64) and compared with the oracle element by element: bucket membership exact, values <= 1e-6 relative as fp32
    and <= 1e-9 as float64."""
    b, n = 64, 16384
    u = np.stack([np.random.RandomState(1234 + i).uniform(size=n) for i in range(b)])
    km = gswm.KeyMaterial.make(KEY, NONCE, gswm.pad_message("lthero", 32), 256)
    du = torch.from_numpy(u).to(cuda_device)
    z32 = gswm.embed_batch_injected(du, (4, 64, 64), km, b, torch.float32).cpu().numpy().reshape(b, n)
    z64 = gswm.embed_batch_injected(du, (4, 64, 64), km, b, torch.float64).cpu().numpy().reshape(b, n)
    for i in range(b):
        ref = O.embed("lthero", KEY, NONCE, u[i], 256)
        assert np.array_equal(z32[i] >= 0, ref >= 0) and np.array_equal(z64[i] >= 0, ref >= 0), i
        assert rel_err(z32[i], ref.astype(np.float32)).max() <= REL_TOL and rel_err(z64[i], ref).max() <= 1e-9, i


def test_embed_injected_edges_and_shared_u(gswm, cuda_device):
    n = 512
    u = np.random.RandomState(5).uniform(size=n)
    u[:4] = [0.0, 1 - 2.0 ** -53, 2.0 ** -53, 0.5]
    km = gswm.KeyMaterial.make(KEY, NONCE, b"abcd", 32)
    z = gswm.embed_batch_injected(torch.from_numpy(u).to(cuda_device), (4, 8, 16), km, 3, torch.float64).cpu().numpy()
    ref = O.embed(b"abcd", KEY, NONCE, u, 32)
    for b in range(3):
        got = z[b].reshape(-1)
        assert np.array_equal(np.isinf(got), np.isinf(ref))
        assert np.array_equal(got >= 0, ref >= 0)
        assert rel_err(got, ref).max() <= 1e-9


# ------------------------------------------------------------------------------------ K3 extract
def _noisy(base, sigma, seed, dtype):
    zn = base.astype(np.float64)
    if sigma:
        zn = zn + sigma * np.random.RandomState(seed).standard_normal(zn.shape)
    return np.clip(zn, -60000.0, 8.0).astype(dtype)


def test_extract_golden_strings(gswm, cuda_device, golden, golden_arrays):
    base = golden_arrays["cli_lthero_z32"]
    msg = gswm.pad_message("lthero", 32)
    for c in golden["extract"]:
        if "noise_seed" not in c:
            continue
        z = _noisy(base, c["sigma"], c["noise_seed"], c["dtype"])
        km = gswm.KeyMaterial.make(KEY, NONCE, msg, 256)
        res = gswm.extract_batch(torch.from_numpy(z).reshape(1, 4, 64, 64).to(cuda_device), km, want_counts=True)
        assert res.bit_strings()[0] == c["extracted_bin"], c["name"]
        assert np.array_equal(res.counts.cpu().numpy()[0].astype(np.uint32), O.vote_counts(z, KEY, NONCE, 256))
        assert int(res.matched[0]) == round(c["bit_accuracy"] * 256)
        assert res.bit_accuracy() == c["bit_accuracy"]


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16, torch.float64])
@pytest.mark.parametrize("shape,L", [((4, 64, 64), 256), ((4, 128, 128), 256), ((4, 64, 64), 32), ((4, 128, 128), 1024),
                                     ((4, 96, 64), 96), ((4, 8, 16), 512), ((4, 64, 64), 2048), ((4, 160, 128), 320),
                                     ((4, 8, 8), 32), ((4, 152, 104), 256), ((4, 72, 72), 64)])
def test_extract_counts_vs_oracle(gswm, cuda_device, dtype, shape, L):
    n = int(np.prod(shape))
    b = 6
    rs = np.random.RandomState(n + L)
    msg = rs.bytes(L // 8)
    km = gswm.KeyMaterial.make(KEY, NONCE, msg, L)
    z = gswm.embed_batch(b, shape, km, 11, 0, 0, cuda_device)
    zn = (z + 1.7 * torch.from_numpy(rs.standard_normal((b, *shape))).to(cuda_device).float()).clamp(max=8.0).to(dtype)
    res = gswm.extract_batch(zn, km, want_counts=True)
    zh = zn.double().cpu().numpy() if dtype == torch.float64 else zn.float().cpu().numpy()
    matched_total = 0
    for i in range(b):
        counts = O.vote_counts(zh[i], KEY, NONCE, L)
        assert np.array_equal(res.counts[i].cpu().numpy().astype(np.uint32), counts), (i, dtype)
        bits = O.recover_message_bits(zh[i], KEY, NONCE, L)
        assert O.bits_to_bytes(bits) == res.messages[i].cpu().numpy().tobytes()
        m = int((bits == np.unpackbits(np.frombuffer(msg, np.uint8))).sum())
        assert int(res.matched[i]) == m
        matched_total += m
    c = res.counters.cpu().numpy()
    assert list(c) == [matched_total, b * L, int(sum(int(x) == L for x in res.matched.cpu())), b, 0, 0]


def test_extract_per_latent_keys(gswm, cuda_device):
    rs = np.random.RandomState(8)
    b, shape, L, n = 33, (4, 64, 64), 256, 16384
    keys = [rs.bytes(32) for _ in range(b)]
    nonces = [rs.bytes(16) for _ in range(b)]
    msgs = [rs.bytes(32) for _ in range(b)]
    km = gswm.KeyMaterial.make(b"".join(keys), b"".join(nonces), b"".join(msgs), L)
    z = gswm.embed_batch(b, shape, km, 3, 0, 0, cuda_device)
    zn = z + 3.0 * torch.randn(z.shape, device=cuda_device, generator=torch.Generator(cuda_device).manual_seed(1))
    zn = zn.clamp(max=8.0)                     # the reference cannot parse cdf(z) * 2 == 2 (z >= 8.29)
    res = gswm.extract_batch(zn, km, want_counts=True)
    zh = zn.cpu().numpy()
    for i in range(b):
        assert np.array_equal(res.counts[i].cpu().numpy().astype(np.uint32), O.vote_counts(zh[i], keys[i], nonces[i], L))
    clean = gswm.extract_batch(z, km)
    assert clean.messages.cpu().numpy().tobytes() == b"".join(msgs)
    assert clean.bit_accuracy() == 1.0


def test_extract_quantiser_edges(gswm, cuda_device, golden):
    """-0.0 and the [-6.957e-17, 0) sliver decode as 1, exactly like int(norm.cdf(z) * 2)."""
    vals = [float(e["z"]) for e in golden["quantise_edges"] if abs(float(e["z"])) < 1e30 or np.isinf(float(e["z"]))]
    z = np.full(512, -1.0, dtype=np.float32)
    z32 = np.array(vals, dtype=np.float32)
    z[:z32.size] = z32
    expect_bits = (z.astype(np.float64) >= O.CDF_HALF_THRESHOLD).astype(np.uint8)
    # cross-check the closed form against scipy on the fp32 values themselves
    assert np.array_equal(O.quantise(z.astype(np.float64)), expect_bits)
    km = gswm.KeyMaterial.make(KEY, NONCE, None, 512)        # one copy: counts == decrypted bits
    res = gswm.extract_batch(torch.from_numpy(z).reshape(1, 4, 8, 16).to(cuda_device), km, want_counts=True)
    assert np.array_equal(res.counts[0].cpu().numpy().astype(np.uint32), O.vote_counts(z, KEY, NONCE, 512))
    # float64 input is compared in float64 on the device: the reference's exact switch-over value and its neighbours
    z64 = np.full(512, -1.0, dtype=np.float64)
    v64 = [float(e["z"]) for e in golden["quantise_edges"] if np.isfinite(float(e["z"])) and float(e["z"]) < 8.0]
    v64 += [O.CDF_HALF_THRESHOLD, np.nextafter(O.CDF_HALF_THRESHOLD, -1.0), np.nextafter(O.CDF_HALF_THRESHOLD, 1.0),
            -6.957291061679417e-17, -6.957291061679418e-17, 5e-324, -5e-324, 1e-300, -1e-300, -7e-17, -6e-17]
    z64[:len(v64)] = v64
    assert np.array_equal(O.quantise(z64), (z64 >= O.CDF_HALF_THRESHOLD).astype(np.uint8))
    res = gswm.extract_batch(torch.from_numpy(z64).reshape(1, 4, 8, 16).to(cuda_device), km, want_counts=True)
    assert np.array_equal(res.counts[0].cpu().numpy().astype(np.uint32), O.vote_counts(z64, KEY, NONCE, 512))
    for dt in (torch.float16, torch.bfloat16):
        zz = torch.tensor([0.0, -0.0, 6e-8, -6e-8, 1.0, -1.0, 8.0, -65504.0] * 64, dtype=dt)
        # bit patterns the arithmetic cannot produce by rounding: smallest / largest subnormals of either sign
        # (bf16 has fp32's range: 0xA4A0 = -6.94e-17 is the last value that still decodes as 1, 0xA4A1 the first 0)
        raw = torch.tensor([0x0001, 0x8001, 0x03FF, 0x83FF, 0x007F, 0x807F, 0x0080, 0x8080, 0xFBFF, 0xA4A0, 0xA4A1,
                            0xA49F, 0xA500, 0x2420], dtype=torch.int32)
        zz[8:8 + raw.numel()] = raw.to(torch.int16).view(dt)
        res = gswm.extract_batch(zz.reshape(1, 4, 8, 16).to(cuda_device), km, want_counts=True)
        assert np.array_equal(res.counts[0].cpu().numpy().astype(np.uint32),
                              O.vote_counts(zz.float().numpy(), KEY, NONCE, 512)), dt


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_extract_saturated_counts_large_latent(gswm, cuda_device, dtype):
    """A 2048 x 2048 image's latent (4 x 256 x 256): R = 1024 copies, a thread sees 256 of them.  With a noise-free
    all-ones message every one of them votes 1 -- the in-register byte-lane counters must not wrap."""
    shape, L = (4, 256, 256), 256
    n = int(np.prod(shape))
    for msg in (b"\xff" * 32, bytes(32), np.random.RandomState(5).bytes(32)):
        km = gswm.KeyMaterial.make(KEY, NONCE, msg, L)
        z = gswm.embed_batch(3, shape, km, 17, 0, 0, cuda_device).to(dtype)
        res = gswm.extract_batch(z, km, want_counts=True)
        want = np.unpackbits(np.frombuffer(msg, np.uint8)).astype(np.uint32) * (n // L)
        for i in range(3):
            assert np.array_equal(res.counts[i].cpu().numpy().astype(np.uint32), want)
        assert res.messages.cpu().numpy().tobytes() == msg * 3


def test_vote_tie_decodes_zero(gswm, cuda_device):
    # R = 2 copies that disagree everywhere -> count 1 of 2 -> strict majority fails -> all zero bits
    L, n = 256, 512
    ks = O.chacha20_keystream(KEY, NONCE, n // 8)
    bits = np.unpackbits(ks)                   # decrypts to all zeros
    bits[L:] ^= 1                              # second copy decrypts to all ones
    z = np.where(bits == 1, 1.0, -1.0).astype(np.float32)
    km = gswm.KeyMaterial.make(KEY, NONCE, bytes(L // 8), L)
    res = gswm.extract_batch(torch.from_numpy(z).reshape(1, 4, 8, 16).to(cuda_device), km, want_counts=True)
    assert (res.counts.cpu().numpy() == 1).all()
    assert res.messages.cpu().numpy().tobytes() == bytes(L // 8)
    assert O.recover_message(z, KEY, NONCE, L) == "0" * L


def test_round_trip_baseline_sizes(gswm, cuda_device):
    """BASELINE config 2/3 size (B = 4096 SD-2.1 latents): every message decodes exactly; config 4 shape."""
    msg = gswm.pad_message("lthero", 32)
    km = gswm.KeyMaterial.make(KEY, NONCE, msg, 256)
    z = gswm.embed_batch(4096, (4, 64, 64), km, 0x5EED, 0, 0, cuda_device)
    res = gswm.extract_batch(z, km)
    assert list(res.counters.cpu().numpy()) == [4096 * 256, 4096 * 256, 4096, 4096, 0, 0]
    # moments of the watermarked noise: it must still look like N(0, 1)
    assert abs(float(z.mean())) < 2e-3 and abs(float(z.std()) - 1.0) < 2e-3
    # sigma = 0.325 regime: ~90 % of signs agree, every message still decodes
    noisy = z + 0.325 * torch.randn(z.shape, device=cuda_device, generator=torch.Generator(cuda_device).manual_seed(99))
    agree = float(((noisy >= 0) == (z >= 0)).float().mean())
    assert 0.89 < agree < 0.91
    res = gswm.extract_batch(noisy, km, want_counts=True)
    assert res.bit_accuracy() == 1.0
    # per-position counts [4096][256] of the WHOLE batch bit-exact against the oracle
    want = O.vote_counts_batch(noisy.cpu().numpy().reshape(4096, -1), KEY, NONCE, 256)
    assert np.array_equal(res.counts.cpu().numpy().astype(np.uint32), want)
    # sigma = 4.3 regime: ~90 % of the DECODED bits survive; counts / messages / matched bits of a
    # subset bit-exact against the oracle, and the batch counters equal the sum of the per-latent figures
    heavy = (z + 4.3 * torch.randn(z.shape, device=cuda_device, generator=torch.Generator(cuda_device).manual_seed(7))).clamp(max=8.0)
    res = gswm.extract_batch(heavy, km, want_counts=True)
    assert 0.88 < res.bit_accuracy() < 0.92
    c = res.counters.cpu().numpy()
    assert int(c[0]) == int(res.matched.sum()) and int(c[1]) == 4096 * 256 and int(c[3]) == 4096
    assert int(c[2]) == int((res.matched == 256).sum())
    ref_bits = np.unpackbits(np.frombuffer(msg, np.uint8))
    want = O.vote_counts_batch(heavy.cpu().numpy().reshape(4096, -1), KEY, NONCE, 256)      # all 4096 latents
    assert np.array_equal(res.counts.cpu().numpy().astype(np.uint32), want)
    want_bits = (want > 64 / 2).astype(np.uint8)                                            # extract.py:99, R = 64
    assert np.array_equal(np.unpackbits(res.messages.cpu().numpy(), axis=1), want_bits)
    assert np.array_equal(res.matched.cpu().numpy(), (want_bits == ref_bits[None, :]).sum(axis=1))
    sub = heavy[:4].cpu().numpy()
    for i in range(4):                                                                      # and the per-latent oracle form agrees
        assert np.array_equal(want[i], O.vote_counts(sub[i], KEY, NONCE, 256))
    del z, noisy, heavy
    zx = gswm.embed_batch(512, (4, 128, 128), km, 1, 0, 0, cuda_device)
    rx = gswm.extract_batch(zx, km)
    assert list(rx.counters.cpu().numpy()) == [512 * 256, 512 * 256, 512, 512, 0, 0]


def test_watermarked_noise_is_standard_normal(gswm, cuda_device):
    """The point of Gaussian Shading: the watermarked latent must be distributed like the N(0,1) noise it replaces.
    67 M elements (4096 SD-2.1 latents) from the in-kernel uniform source (Philox4x32-7, 23-bit grid), all statistics
    within 5 sigma of their expectation.
    Shared key (every latent carries the same bucket pattern, as in the reference): inside each bucket |z|'s uniform is
    chi-square-flat over 1000 bins, even moments match, magnitudes are uncorrelated along and across latents.
    One key per latent (independent bucket patterns): Phi(z) is chi-square-flat, odd moments vanish, z is serially
    uncorrelated."""
    rs = np.random.RandomState(31337)
    B, n, bins = 4096, 16384, 1000
    N = B * n
    five_sigma_chi2 = 5 * np.sqrt(2 * (bins - 1))

    def corr(a, b):
        a, b = a.double().reshape(-1), b.double().reshape(-1)
        a, b = a - a.mean(), b - b.mean()
        return float((a * b).mean() / (a.std() * b.std()))

    def chi2_flat(x):
        h = torch.histc(x, bins=bins, min=0.0, max=1.0).cpu().numpy()
        return float(((h - x.numel() / bins) ** 2 / (x.numel() / bins)).sum())

    lim = 5 / np.sqrt(N - B)
    # ---- shared key: the throughput kernel ----
    km = gswm.KeyMaterial.make(KEY, NONCE, rs.bytes(32), 256)
    z = gswm.embed_batch(B, (4, 64, 64), km, 0xC0FFEE, 0, 0, cuda_device).reshape(B, n)
    v = (2.0 * torch.special.ndtr(z.double()) - 1.0).abs()            # the in-bucket uniform
    for sign in (z >= 0, z < 0):
        c = chi2_flat(v[sign])
        assert abs(c - (bins - 1)) < five_sigma_chi2, c
    zd = z.double()
    assert abs(float((zd ** 2).mean()) - 1) < 5 * np.sqrt(2 / N) and abs(float((zd ** 4).mean()) - 3) < 5 * np.sqrt(96 / N)
    za = z.abs()
    for lag in (1, 4, 256, 1024):                                     # neighbours, same lane, same thread, next super-iteration
        assert abs(corr(za[:, lag:], za[:, :-lag])) < lim, lag
    assert abs(corr(za[1:], za[:-1])) < lim                           # the same element of consecutive latents
    del z, v, zd, za
    # ---- one key / nonce / message per latent ----
    kmp = gswm.KeyMaterial.make(rs.bytes(32 * B), rs.bytes(16 * B), rs.bytes(32 * B), 256)
    z = gswm.embed_batch(B, (4, 64, 64), kmp, 0xBEEF, 0, 0, cuda_device).reshape(B, n)
    c = chi2_flat(torch.special.ndtr(z.double()))
    assert abs(c - (bins - 1)) < five_sigma_chi2, c
    zd = z.double()
    assert abs(float(zd.mean())) < 5 / np.sqrt(N) and abs(float((zd ** 2).mean()) - 1) < 5 * np.sqrt(2 / N)
    assert abs(float((zd ** 3).mean())) < 5 * np.sqrt(15 / N) and abs(float((zd ** 4).mean()) - 3) < 5 * np.sqrt(96 / N)
    assert abs(corr(z[:, 1:], z[:, :-1])) < lim and abs(corr(z[1:], z[:-1])) < lim


def test_coscheduled_step_matches_separate_calls(gswm, cuda_device):
    """embed_extract_batch (embed and extract on two streams, sharing the SMs) returns exactly what the two calls
    return one after the other -- including when it is called repeatedly with the results consumed at once."""
    msg = gswm.pad_message("lthero", 32)
    km = gswm.KeyMaterial.make(KEY, NONCE, msg, 256)
    z0 = gswm.embed_batch(1024, (4, 64, 64), km, 11, 0, 0, cuda_device)
    noisy = z0 + 0.8 * torch.randn(z0.shape, device=cuda_device, generator=torch.Generator(cuda_device).manual_seed(3))
    want_z = gswm.embed_batch(1024, (4, 64, 64), km, 12, 0, 5000, cuda_device)
    want = gswm.extract_batch(noisy, km, want_counts=True)
    for _ in range(4):
        z, res = gswm.embed_extract_batch(1024, (4, 64, 64), km, 12, noisy, first_latent=5000, want_counts=True)
        assert torch.equal(z, want_z)
        assert torch.equal(res.messages, want.messages) and torch.equal(res.counts, want.counts)
        assert torch.equal(res.matched, want.matched) and torch.equal(res.counters, want.counters)


def test_back_to_back_jobs_with_changing_keys(gswm, cuda_device):
    """Programmatic dependent launch lets a kernel's prologue (its ChaCha20 keystream) run under the previous kernel's
    tail.  60 jobs with fresh key material, alternating shapes and both key modes, enqueued without any host sync,
    must each decode to their own message and reproduce the oracle's bucket bits."""
    rs = np.random.RandomState(4242)
    shapes = [((4, 64, 64), 256), ((4, 128, 128), 256), ((4, 8, 16), 32), ((4, 96, 64), 96)]
    jobs = []
    for it in range(60):
        shape, L = shapes[it % len(shapes)]
        b = int(rs.randint(1, 700))
        per = it % 5 == 4
        rows = b if per else 1
        key, nonce, msg = rs.bytes(32 * rows), rs.bytes(16 * rows), rs.bytes((L // 8) * rows)
        km = gswm.KeyMaterial.make(key, nonce, msg, L)
        z = gswm.embed_batch(b, shape, km, it, 0, 0, cuda_device)
        res = gswm.extract_batch(z, km)
        jobs.append((shape, L, b, per, key, nonce, msg, z[:1].clone(), res))
    torch.cuda.synchronize()
    for shape, L, b, per, key, nonce, msg, z0, res in jobs:
        want = msg if per else msg * b
        assert res.messages.cpu().numpy().tobytes() == want
        assert list(res.counters.cpu().numpy()) == [b * L, b * L, b, b, 0, 0]
        n = int(np.prod(shape))
        y = O.bucket_bits(O.frame_message(msg[:L // 8], n, L)[1], key[:32], nonce[:16])[:n]
        assert np.array_equal(z0.cpu().numpy().reshape(-1) >= 0, y == 1)


def _sign_checksum(z):
    """Order-sensitive 64-bit checksum of a batch's bucket bits (sum over elements of sign * odd weight), on the GPU."""
    n = z[0].numel()
    w = (torch.arange(n, device=z.device, dtype=torch.int64) * 2654435761 + 1) & 0xFFFFFFFF
    return ((z.reshape(z.shape[0], n) >= 0).to(torch.int64) * w).sum(dim=1)


def test_config4_sdxl_full_size(gswm, cuda_device):
    """BASELINE config 4 at full size: 65 536 SDXL latents (4x128x128, 17.2 GB), embed + extract, checked through
    size-independent properties: every message decodes exactly, bucket bits (a function of key/nonce/message only)
    are identical for every latent and equal the oracle's, a 256-latent subset matches the oracle element-wise."""
    free, _ = torch.cuda.mem_get_info(cuda_device)
    B, shape, n = 65536, (4, 128, 128), 65536
    if free < B * n * 4 + (8 << 30):
        pytest.skip("needs ~26 GB of free HBM")
    msg = gswm.pad_message("lthero", 32)
    km = gswm.KeyMaterial.make(KEY, NONCE, msg, 256)
    z = gswm.embed_batch(B, shape, km, 0x5EED, 0, 0, cuda_device)
    res = gswm.extract_batch(z, km)
    assert list(res.counters.cpu().numpy()) == [B * 256, B * 256, B, B, 0, 0]
    assert bool((res.matched == 256).all())
    # checksum of checksums: the sign pattern is the same for all latents and equals the oracle's
    y = O.bucket_bits(O.frame_message("lthero", n, 256)[1], KEY, NONCE)[:n]
    w = (np.arange(n, dtype=np.int64) * 2654435761 + 1) & 0xFFFFFFFF
    expect = int((y.astype(np.int64) * w).sum())
    cs = torch.cat([_sign_checksum(z[i:i + 4096]) for i in range(0, B, 4096)])
    assert bool((cs == expect).all())
    # sharding transparency at full size: latents [B-256, B) of the big batch == a 256-latent job at first_latent = B-256
    tail = gswm.embed_batch(256, shape, km, 0x5EED, 0, B - 256, cuda_device)
    assert torch.equal(tail, z[B - 256:])
    sub = tail.cpu().numpy().reshape(256, n)
    for i in range(256):                                    # a 256-sample subset, element by element
        ref = O.embed_gswm("lthero", KEY, NONCE, 0x5EED, 0, B - 256 + i, n, 256)
        assert np.array_equal(sub[i] >= 0, ref >= 0) and rel_err(sub[i], ref).max() <= REL_TOL, i
    noisy = (tail + 1.5 * torch.randn(tail.shape, device=cuda_device, generator=torch.Generator(cuda_device).manual_seed(4))).clamp(max=8.0)
    rs = gswm.extract_batch(noisy, km, want_counts=True)
    assert np.array_equal(rs.counts.cpu().numpy().astype(np.uint32),
                          O.vote_counts_batch(noisy.cpu().numpy().reshape(256, n), KEY, NONCE, 256))
    # no two latents share their noise: |z| differs between latents (same signs, different magnitudes)
    assert not torch.equal(z[0].abs(), z[1].abs()) and not torch.equal(z[0].abs(), z[B - 1].abs())


def test_config5_per_latent_keys_1m(gswm, cuda_device):
    """BASELINE config 5 at full size: 1 M SD-2.1 latents, each with its own key / nonce / message (RandomState(2025)),
    streamed through the device API in 8 chunks of 131 072 (8.6 GB each).  Every decoded message must equal the
    message embedded in that latent; a ~1024-latent subset is checked against the oracle through `cryptography`."""
    free, _ = torch.cuda.mem_get_info(cuda_device)
    total, chunk, shape, n, L = 1 << 20, 1 << 17, (4, 64, 64), 16384, 256
    if free < chunk * n * 4 + (4 << 30):
        pytest.skip("needs ~13 GB of free HBM")
    rs = np.random.RandomState(2025)
    keys = np.frombuffer(rs.bytes(32 * total), np.uint8).reshape(total, 32)
    nonces = np.frombuffer(rs.bytes(16 * total), np.uint8).reshape(total, 16)
    msgs = np.frombuffer(rs.bytes(32 * total), np.uint8).reshape(total, 32)
    out = torch.empty((chunk, *shape), dtype=torch.float32, device=cuda_device)
    counters = torch.zeros(gswm._lib.N_COUNTERS, dtype=torch.int64, device=cuda_device)
    for c in range(total // chunk):
        sl = slice(c * chunk, (c + 1) * chunk)
        km = gswm.KeyMaterial.make(keys[sl], nonces[sl], msgs[sl], L)
        gswm.embed_batch(chunk, shape, km, 2025, 0, c * chunk, cuda_device, out=out)
        res = gswm.extract_batch(out, km, counters=counters)
        assert torch.equal(res.messages.cpu(), torch.from_numpy(msgs[sl].copy()))
        if c in (0, 7):                                     # 2 x 512 = 1024 samples through `cryptography`
            idx = sorted({0, 1, chunk // 2, chunk - 1} | set(np.random.RandomState(c).randint(0, chunk, size=508).tolist()))
            zs = out[idx].cpu().numpy().reshape(len(idx), n)
            for j, i in enumerate(idx):
                g = c * chunk + i
                k, no, m = keys[g].tobytes(), nonces[g].tobytes(), msgs[g].tobytes()
                y = O.bucket_bits(O.frame_message(m, n, L)[1], k, no, keystream=O.chacha20_keystream_lib)[:n]
                assert np.array_equal(zs[j] >= 0, y == 1)
                ref = O.embed_gswm(m, k, no, 2025, 0, g, n, L)
                assert rel_err(zs[j], ref).max() <= REL_TOL
    assert list(counters.cpu().numpy()) == [total * L, total * L, total, total, 0, 0]


def test_c_host_round_trip(gswm, cuda_device, tmp_path):
    """tests/c_abi/roundtrip.c: a C program on the CUDA runtime alone (no Python, no torch) embeds, decodes and scores a batch
    through libgswm.so, device API and host-buffer pipe, and checks known answers of the default key."""
    import subprocess
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    exe = tmp_path / "roundtrip"
    libdir = os.path.dirname(gswm._lib.LIB_PATH)
    cc = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", f"-I{os.path.join(ROOT, 'include')}", f"-I{cuda}/include",
                         os.path.join(ROOT, "tests", "c_abi", "roundtrip.c"), "-o", str(exe), f"-L{libdir}", "-lgswm",
                         f"-L{cuda}/lib64", "-lcudart", f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{cuda}/lib64"],
                        capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0 and "roundtrip ok" in run.stdout, run.stdout + run.stderr


# ------------------------------------------------------------------------------------ host-buffer pipe
def test_host_pipe_matches_device_api(gswm, cuda_device):
    msg = gswm.pad_message("lthero", 32)
    km = gswm.KeyMaterial.make(KEY, NONCE, msg, 256)
    pipe = gswm.HostPipe(0, max_elems=16384, chunk_latents=8)
    out = torch.empty((27, 4, 64, 64), dtype=torch.float32).pin_memory()
    pipe.embed(out, km, 5, 0, 100)
    dev = gswm.embed_batch(27, (4, 64, 64), km, 5, 0, 100, cuda_device)
    assert torch.equal(out, dev.cpu())
    noisy = (out + 2.0 * torch.randn(out.shape, generator=torch.Generator().manual_seed(3)))
    msgs, cnt, matched, counters, flags = pipe.extract(noisy, km, want_counts=True)
    r = gswm.extract_batch(noisy.to(cuda_device), km, want_counts=True)
    assert np.array_equal(msgs, r.messages.cpu().numpy())
    assert np.array_equal(cnt, r.counts.cpu().numpy())
    assert np.array_equal(matched, r.matched.cpu().numpy())
    assert np.array_equal(counters, r.counters.cpu().numpy())
    assert np.array_equal(flags, r.flags.cpu().numpy()) and flags.any()    # sigma = 2 noise: most latents hold a z >= 8.29
    for i in range(27):                                                    # ... exactly the ones the reference refuses
        rejected = bool((noisy[i].double().numpy() >= O.CDF_ONE_THRESHOLD).any())
        assert flags[i] == (gswm._lib.FLAG_RANGE if rejected else 0)
    # fp16 host input, per-latent keys
    rs = np.random.RandomState(4)
    kmp = gswm.KeyMaterial.make(rs.bytes(32 * 27), rs.bytes(16 * 27), rs.bytes(32 * 27), 256)
    pipe.embed(out, kmp, 6)
    m2, _, mt2, c2, _ = pipe.extract(out.half(), kmp)
    assert m2.tobytes() == kmp.msgs.tobytes() and list(c2) == [27 * 256, 27 * 256, 27, 27, 0, 0]
    u = np.random.RandomState(1234).uniform(size=16384)
    zi = pipe.embed_injected(u, (4, 64, 64), km, 1, np.float64)
    assert rel_err(zi.reshape(-1), O.embed("lthero", KEY, NONCE, u, 256)).max() <= 1e-9
    pipe.close()


# ------------------------------------------------------------------------------------ reference-named drop-ins
def test_dropin_gs_insert_and_extract(gswm, cuda_device, golden, golden_arrays, tmp_path, monkeypatch):
    from gswm import extract as gx
    from gswm import gs_insert as gi

    monkeypatch.chdir(tmp_path)
    opt = types.SimpleNamespace(key_hex=O.DEFAULT_KEY_HEX, nonce_hex=O.DEFAULT_NONCE_HEX)
    np.random.seed(1234)                      # same stream as RandomState(1234) used for the golden run
    z = gi.gs_watermark_init_noise(opt, "lthero")
    assert z.shape == (4, 64, 64) and z.dtype == np.float64
    assert rel_err(z.reshape(-1)[:512], golden_arrays["cli_lthero_z64_head"]).max() <= 1e-9
    assert np.array_equal(z.astype(np.float32) >= 0, golden_arrays["cli_lthero_z32"] >= 0)
    lines = (tmp_path / "info_data.txt").read_text().splitlines()
    assert lines[1:] == golden["info_data_cli_tail"][1:] and lines[0].startswith("Time: ")
    opt.nonce_hex = ""                        # nonce falls back to key_hex[16:48]
    np.random.seed(1236)
    z2 = gi.gs_watermark_init_noise(opt, "lthero")
    c = [e for e in golden["embed_cli"] if e["name"] == "cli_nonce_fallback"][0]
    assert sha(np.packbits((z2.reshape(-1) >= 0).astype(np.uint8))) == c["sha256_signs"]

    args = types.SimpleNamespace(key=KEY, nonce=NONCE, l=1, message_length=256)
    base = golden_arrays["cli_lthero_z32"]
    for c in golden["extract"]:
        if "noise_seed" not in c:
            continue
        zt = torch.from_numpy(_noisy(base, c["sigma"], c["noise_seed"], c["dtype"])).reshape(1, 4, 64, 64)
        got = gx.recover_exactracted_message(zt, args)
        assert got == c["extracted_bin"], c["name"]
        orig, acc = gx.calculate_bit_accuracy((b"lthero" + bytes(26)).hex(), got)
        assert acc == c["bit_accuracy"] and orig == c["original_bin"]
    for bad in golden["quantise_raises"]:
        t = torch.full((1, 1, 16, 32), 0.5, dtype=torch.float64)
        t[0, 0, 0, 0] = float(bad["z"])
        with pytest.raises(ValueError):
            gx.recover_exactracted_message(t, types.SimpleNamespace(key=KEY, nonce=NONCE, l=1, message_length=512))
    batch = gi.gs_watermark_init_noise_batch(opt, "lthero", n_samples=5, seed=1)
    assert batch.shape == (5, 4, 64, 64) and batch.is_cuda


def test_dropin_comfy_and_webui(gswm, cuda_device, golden, golden_arrays, tmp_path, monkeypatch):
    """ComfyUI node function / GSLatent node and both webui scripts against the vectors the reference produced."""
    import io

    from gswm import comfy_nodes as cn
    from gswm import extract as gx
    from gswm import webui_v152 as w5
    from gswm import webui_v160 as w6

    monkeypatch.chdir(tmp_path)
    for c in golden["embed_comfy"]:
        z = cn.gs_watermark_init_noise(O.DEFAULT_KEY_HEX, O.DEFAULT_NONCE_HEX, "cpu", c["message"], 1, c["seed"], c["width"],
                                       c["height"], c["message_length"])
        assert z.dtype == torch.float32 and not z.is_cuda and list(z.shape) == c["shape"], c["name"]
        zz = z.numpy().reshape(-1)
        assert sha(np.packbits((zz >= 0).astype(np.uint8))) == c["sha256_signs"], c["name"]
        assert rel_err(zz[:256], golden_arrays[c["name"] + "_z32_head"]).max() <= REL_TOL, c["name"]
    lines = (tmp_path / "info_data.txt").read_text().splitlines()
    assert lines[-9:] == golden["info_data_comfy_tail"][1:] and lines[-10].startswith("Time: ")
    lat, first = cn.GSLatent().create_gs_latents(O.DEFAULT_KEY_HEX, O.DEFAULT_NONCE_HEX, "lthero", 3, 1, 42, 512, 512, 256)
    assert list(lat["samples"].shape) == golden["gslatent_seeded"]["shape"]
    assert torch.equal(lat["samples"][0], lat["samples"][2]) and torch.equal(first, lat["samples"][0])
    np.random.seed(5)
    lat2, _ = cn.GSLatent().create_gs_latents(O.DEFAULT_KEY_HEX, O.DEFAULT_NONCE_HEX, "lthero", 3, 0, 0, 512, 512, 256)
    ref_u = np.random.RandomState(5).uniform(size=(3, 16384))
    for b in range(3):
        ref = O.embed("lthero", KEY, NONCE, ref_u[b], 256)
        assert rel_err(lat2["samples"][b].numpy().reshape(-1), ref).max() <= REL_TOL
    assert set(cn.NODE_CLASS_MAPPINGS) == {"Lthero_GSLatent", "Lthero_GS_KSamplerAdvanced"}

    for c in golden["embed_webui"]:
        for mod in (w5, w6):
            w5.global_message, w5.global_key, w5.global_nonce = c["message"], O.DEFAULT_KEY_HEX, O.DEFAULT_NONCE_HEX
            w5.global_use_randomSeed, w5.global_randomSeed, w5.global_use_repeat = 1, c["seed"], c["use_repeat"]
            z = mod.init_gs_Z_s_T()
            assert z.shape == (4, 64, 64) and z.dtype == np.float64
            assert sha(np.packbits((z.reshape(-1) >= 0).astype(np.uint8))) == c["sha256_signs"], c["name"]
            assert rel_err(z.reshape(-1)[:256], golden_arrays[c["name"] + "_z64_head"]).max() <= 1e-9
    lines = (tmp_path / "info_data.txt").read_text().splitlines()
    assert lines[-5:] == golden["info_data_webui_tail"][1:]
    t = w5.advanced_creator((1, 4, 64, 64), [1])
    assert t.shape == (1, 4, 64, 64) and t.dtype == torch.float32 and t.is_cuda
    assert w6.global_randomSeed == w5.global_randomSeed

    # batched evaluation front-end (extract.py:134-175)
    base = torch.from_numpy(golden_arrays["cli_lthero_z32"])
    batch = torch.stack([base, base + 4.3 * torch.randn(base.shape, generator=torch.Generator().manual_seed(1)),
                         (base + 0.5 * torch.randn(base.shape, generator=torch.Generator().manual_seed(2)))]).clamp(max=8.0)
    args = types.SimpleNamespace(key=KEY, nonce=NONCE, l=1, message_length=256, original_message_hex=(b"lthero" + bytes(26)).hex(),
                                 key_hex=O.DEFAULT_KEY_HEX, nonce_hex=O.DEFAULT_NONCE_HEX, num_inference_steps=30, scheduler="DDIM")
    buf = io.StringIO()
    gx.write_batch_info(buf, args)
    strings, accs, avg = gx.evaluate_latents(["a.png", "b.png", "c.png"], batch.half(), args, buf)
    for i in range(3):
        assert strings[i] == O.recover_message(batch[i].half().numpy(), KEY, NONCE, 256)
        assert accs[i] == O.calculate_bit_accuracy(args.original_message_hex, strings[i])[1]
    text = buf.getvalue().splitlines()
    assert text[0] == "=" * 40 + "Batch Info" + "=" * 40 and text[2] == f"key_hex,{O.DEFAULT_KEY_HEX}"
    assert text[8] == f"a.png, Bit Accuracy, {accs[0]}" and text[11] == f"Average Bit Accuracy, {avg}"
    assert accs[0] == 1.0 and accs[2] == 1.0 and accs[1] < 1.0


def _golden_fake_urandom():
    """The deterministic os.urandom stand-in of the golden run (one definition: tests/golden/make_golden.py)."""
    import importlib.util
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_golden.py")
    spec = importlib.util.spec_from_file_location("make_golden", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)                 # defines helpers only; main() is what touches /root/reference
    return mod._FakeUrandom()


def test_dropin_comfy_unseeded_batch_golden(gswm, cuda_device, golden, golden_arrays, tmp_path, monkeypatch):
    """GSLatent with use_seed=0 (nodes.py:236-237): batch_size independent calls.  Against the reference's own run:
    latents (numpy global stream), one info_data.txt record per latent with the widget's seed logged, and -- with an empty
    message and key -- a fresh os.urandom message / key / nonce PER LATENT in the reference's call order."""
    import os

    from gswm import comfy_nodes as cn

    monkeypatch.chdir(tmp_path)
    g = golden["gslatent_unseeded"]
    np.random.seed(g["np_seed"])
    lat, first = cn.GSLatent().create_gs_latents(O.DEFAULT_KEY_HEX, O.DEFAULT_NONCE_HEX, "lthero", g["batch_size"], 0,
                                                 g["widget_seed"], 512, 512, 256)
    s = lat["samples"].numpy()
    assert list(s.shape) == g["shape"] and s.dtype == np.float32 and torch.equal(first, lat["samples"][0])
    for i in range(g["batch_size"]):
        assert sha(np.packbits((s[i].reshape(-1) >= 0).astype(np.uint8))) == g["sha256_signs_each"][i]
    assert rel_err(s.reshape(g["batch_size"], -1)[:, :128], golden_arrays["gslatent_unseeded_heads"]).max() <= REL_TOL
    lines = (tmp_path / "info_data.txt").read_text().splitlines()
    want = g["info_data_tail"]
    assert len(lines) == len(want) == 10 * g["batch_size"]
    for a, b in zip(lines, want):
        assert a == b or (a.startswith("Time: ") and b.startswith("Time: "))
    assert f"randomSeed: {g['widget_seed']}" in lines

    g = golden["gslatent_unseeded_random"]
    (tmp_path / "info_data.txt").unlink()
    monkeypatch.setattr(os, "urandom", _golden_fake_urandom())
    np.random.seed(g["np_seed"])
    lat, _ = cn.GSLatent().create_gs_latents("", "", "", g["batch_size"], 0, g["widget_seed"], 256, 256, -1)
    monkeypatch.undo()
    monkeypatch.chdir(tmp_path)
    s = lat["samples"].numpy()
    assert list(s.shape) == g["shape"]
    for i in range(g["batch_size"]):
        assert sha(np.packbits((s[i].reshape(-1) >= 0).astype(np.uint8))) == g["sha256_signs_each"][i]
    assert rel_err(s.reshape(g["batch_size"], -1)[:, :128], golden_arrays["gslatent_unseeded_random_heads"]).max() <= REL_TOL
    lines = (tmp_path / "info_data.txt").read_text().splitlines()
    want = g["info_data_tail"]
    assert len(lines) == len(want)
    for a, b in zip(lines, want):
        assert a == b or (a.startswith("Time: ") and b.startswith("Time: "))
    assert lines[1] != lines[11] and lines[3] != lines[13]          # per-latent keys and messages


def test_argument_errors(gswm, cuda_device):
    km = gswm.KeyMaterial.make(KEY, NONCE, bytes(32), 256)
    with pytest.raises(ValueError):
        gswm.embed_batch(1, (3, 5, 5), km, 0, device=cuda_device)          # not a multiple of 4
    with pytest.raises(ValueError):
        gswm.extract_batch(torch.zeros((1, 4, 96, 64), device=cuda_device), gswm.KeyMaterial.make(KEY, NONCE, None, 640))
    lib = gswm._lib.lib()
    job = gswm._lib.Job(1, 16384, 256, 0, 1, 1, 1)
    assert lib.gswm_embed(C.byref(job), 0, 0, 0, 16, None) == -7                  # misaligned key pointer
    job = gswm._lib.Job(1, 16384, 250, 0, 16, 16, 16)
    assert lib.gswm_embed(C.byref(job), 0, 0, 0, 16, None) == -3
    job = gswm._lib.Job(1, 1002, 32, 0, 16, 16, 16)
    assert lib.gswm_embed(C.byref(job), 0, 0, 0, 16, None) == -2
    assert lib.gswm_embed(None, 0, 0, 0, 16, None) == -1
    assert "multiple of 4" in gswm._lib.strerror(-2)


def test_empty_batches(gswm, cuda_device):
    """Zero latents is a valid job everywhere: nothing is launched, outputs are empty, counters stay zero."""
    km = gswm.KeyMaterial.make(KEY, NONCE, gswm.pad_message("lthero", 32), 256)
    before = gswm.launch_count()
    z = gswm.embed_batch(0, (4, 64, 64), km, 1, device=cuda_device)
    assert tuple(z.shape) == (0, 4, 64, 64) and z.dtype == torch.float32 and z.is_cuda
    res = gswm.extract_batch(z, km, want_counts=True)
    assert tuple(res.messages.shape) == (0, 32) and tuple(res.counts.shape) == (0, 256) and tuple(res.matched.shape) == (0,)
    assert res.counters.cpu().tolist() == [0] * 6 and res.bit_strings() == [] and np.isnan(res.bit_accuracy())
    pipe = gswm.HostPipe(cuda_device.index or 0, max_elems=16384, chunk_latents=4)
    out = pipe.embed(np.empty((0, 4, 64, 64), dtype=np.float32), km, 1)
    msgs, cnt, matched, counters, flags = pipe.extract(out, km, want_counts=True)
    assert msgs.shape == (0, 32) and cnt.shape == (0, 256) and matched.shape == (0,) and counters.tolist() == [0] * 6 and flags.shape == (0,)
    pipe.close()
    assert gswm.launch_count() == before
    # the C ABI itself: n_latents == 0 is success, before any launch
    lib = gswm._lib.lib()
    flat = torch.zeros(96, dtype=torch.uint8, device=cuda_device)
    buf = torch.zeros(64, dtype=torch.float32, device=cuda_device)
    job = gswm._lib.Job(0, 16384, 256, 0, flat.data_ptr(), flat.data_ptr() + 32, flat.data_ptr() + 48)
    assert lib.gswm_embed(C.byref(job), 0, 0, 0, buf.data_ptr(), None) == 0
    assert lib.gswm_extract(C.byref(job), buf.data_ptr(), 0, flat.data_ptr(), None, None, None, None, None) == 0
    assert gswm.launch_count() == before


@pytest.mark.parametrize("shape,L", [((4, 96, 64), 96), ((4, 8, 16), 32), ((4, 152, 104), 256), ((4, 128, 128), 1024)])
def test_per_latent_keys_on_ragged_and_multi_tile_latents(gswm, cuda_device, shape, L):
    """Distinct key / nonce / message per latent on shapes that are not one whole tile: partial last tile, less than one
    ChaCha block row, several tiles -- embed against the oracle, extract counts and messages against the oracle."""
    rs = np.random.RandomState(int(np.prod(shape)) + L)
    b, n = 9, int(np.prod(shape))
    keys = [rs.bytes(32) for _ in range(b)]
    nonces = [rs.bytes(16) for _ in range(b)]
    msgs = [rs.bytes(L // 8) for _ in range(b)]
    km = gswm.KeyMaterial.make(b"".join(keys), b"".join(nonces), b"".join(msgs), L)
    z = gswm.embed_batch(b, shape, km, 5, 0, 40, cuda_device)
    zh = z.cpu().numpy().reshape(b, n)
    ref = oracle_embed_batch(msgs, keys, nonces, n, L, 5, 0, 40, b)
    assert np.array_equal(zh >= 0, ref >= 0)
    assert rel_err(zh, ref).max() <= REL_TOL
    if n % L == 0:
        noisy = (z + 2.0 * torch.randn(z.shape, device=cuda_device, generator=torch.Generator(cuda_device).manual_seed(3))).clamp(max=8.0)
        res = gswm.extract_batch(noisy, km, want_counts=True)
        nh = noisy.cpu().numpy().reshape(b, n)
        for i in range(b):
            assert np.array_equal(res.counts[i].cpu().numpy().astype(np.uint32), O.vote_counts(nh[i], keys[i], nonces[i], L)), i
            assert O.bits_to_bytes(O.recover_message_bits(nh[i], keys[i], nonces[i], L)) == res.messages[i].cpu().numpy().tobytes()
        assert gswm.extract_batch(z, km).messages.cpu().numpy().tobytes() == b"".join(msgs)


def test_large_latent_index_seed_offset_and_longest_message(gswm, cuda_device):
    """64-bit corners of the uniform source (global latent index 2^40, full-width seed, offset just below 2^62) against the
    oracle, and the longest message the extract kernel takes (8192 bits); one more bit is a range error, not a launch."""
    shape, n, L = (4, 128, 128), 65536, 256
    msg = gswm.pad_message("lthero", 32)
    km = gswm.KeyMaterial.make(KEY, NONCE, msg, L)
    seed, offset, first = 0xFEDCBA9876543210, (1 << 62) - 5, (1 << 40) + 7
    z = gswm.embed_batch(2, shape, km, seed, offset, first, cuda_device).cpu().numpy().reshape(2, n)
    for i in range(2):
        ref = O.embed_gswm("lthero", KEY, NONCE, seed, offset, first + i, n, L)
        assert np.array_equal(z[i] >= 0, ref >= 0) and rel_err(z[i], ref).max() <= REL_TOL
    with pytest.raises(gswm.GswmError):
        gswm.embed_batch(1, shape, km, seed, 1 << 62, 0, cuda_device)      # offset >= 2^62 is out of range

    L = 8192
    rs = np.random.RandomState(8192)
    big = rs.bytes(L // 8)
    km = gswm.KeyMaterial.make(KEY, NONCE, big, L)
    zz = gswm.embed_batch(3, shape, km, 1, 0, 0, cuda_device)
    noisy = (zz + 0.8 * torch.randn(zz.shape, device=cuda_device, generator=torch.Generator(cuda_device).manual_seed(9))).clamp(max=8.0)
    res = gswm.extract_batch(noisy, km, want_counts=True)
    nh = noisy.cpu().numpy().reshape(3, n)
    for i in range(3):
        assert np.array_equal(res.counts[i].cpu().numpy().astype(np.uint32), O.vote_counts(nh[i], KEY, NONCE, L))
    assert gswm.extract_batch(zz, km).messages.cpu().numpy().tobytes() == big * 3
    with pytest.raises(gswm.GswmError):
        gswm.extract_batch(torch.zeros((1, 4, 128, 128), device=cuda_device), gswm.KeyMaterial.make(KEY, NONCE, None, 16384))


def test_counter_word3_changes_inside_a_launch(gswm, cuda_device):
    """Uniforms v4: the Philox counter carries T = (latent * tiles + tile) * 4 + super-iteration split over two words,
    (T & 0xFFFFFF) next to the lane index and T >> 24 next to the call index.  The lane-independent half of the rounds is
    tabulated per CTA from the T >> 24 word, so a launch that straddles a multiple of 2^24 must switch table entries in the
    middle of a CTA's latent loop -- shared key (persistent grid, table per interval) and per-latent keys (table per latent)."""
    shape, n, L = (4, 64, 64), 16384, 256
    cross = 1 << 22                                                      # tiles = 1: T = 4 * latent crosses 2^24 here
    first, b = cross - 300, 600
    km = gswm.KeyMaterial.make(KEY, NONCE, gswm.pad_message("lthero", 32), L)
    z = gswm.embed_batch(b, shape, km, 0x5EED, 0, first, cuda_device)
    pick = [0, 1, 298, 299, 300, 301, 302, 599] + np.random.RandomState(5).randint(0, b, size=8).tolist()
    zh = z[pick].cpu().numpy().reshape(len(pick), n)
    for j, i in enumerate(pick):
        ref = O.embed_gswm("lthero", KEY, NONCE, 0x5EED, 0, first + i, n, L)
        assert np.array_equal(zh[j] >= 0, ref >= 0) and rel_err(zh[j], ref).max() <= REL_TOL, i
    # the same latents from launches that start elsewhere (other CTA / interval mapping): bit-identical
    assert torch.equal(gswm.embed_batch(7, shape, km, 0x5EED, 0, cross - 3, cuda_device), z[297:304])
    rs = np.random.RandomState(77)
    kb, nb, mb = rs.bytes(32 * 6), rs.bytes(16 * 6), rs.bytes(32 * 6)
    kmp = gswm.KeyMaterial.make(kb, nb, mb, L)
    zp = gswm.embed_batch(6, shape, kmp, 9, 0, cross - 3, cuda_device).cpu().numpy().reshape(6, n)
    for i in range(6):
        ref = O.embed_gswm(mb[32 * i:32 * i + 32], kb[32 * i:32 * i + 32], nb[16 * i:16 * i + 16], 9, 0, cross - 3 + i, n, L)
        assert np.array_equal(zp[i] >= 0, ref >= 0) and rel_err(zp[i], ref).max() <= REL_TOL, i
    # the counter holds 54 bits of T: the last latent that fits, and one more (a range error, not a wrapped counter)
    last = (1 << 52) - 1
    zl = gswm.embed_batch(1, shape, km, 1, 0, last, cuda_device).cpu().numpy().reshape(n)
    ref = O.embed_gswm("lthero", KEY, NONCE, 1, 0, last, n, L)
    assert np.array_equal(zl >= 0, ref >= 0) and rel_err(zl, ref).max() <= REL_TOL
    with pytest.raises(gswm.GswmError):
        gswm.embed_batch(2, shape, km, 1, 0, last, cuda_device)


# ------------------------------------------------------------------------------------ uniforms v3 / v4, generator pins
def test_top_cell_refinement_matches_oracle(gswm, cuda_device):
    """Uniforms v3: an element whose 23-bit m falls in the outermost cell (probability 2^-23) is refined by 28 more Philox
    bits, so |z| can reach 8.21 like the reference's 53-bit uniforms.  Latents 404, 812, 1399 and 1458 of seed 0x5EED each
    hold one such element (found by scanning the oracle's integers); the kernel must reproduce the oracle there too."""
    km = gswm.KeyMaterial.make(KEY, NONCE, gswm.pad_message("lthero", 32), 256)
    n = 16384
    for latent, elem in ((404, 503), (812, 4346), (1399, 14014), (1458, 11327)):
        m = O.gswm_uniform_ints(0x5EED, 0, latent, n)
        assert m[elem] == O.GSWM_TOP_CELL and (m == O.GSWM_TOP_CELL).sum() == 1
        z = gswm.embed_batch(1, (4, 64, 64), km, 0x5EED, 0, latent, cuda_device).cpu().numpy().reshape(-1)
        ref = O.embed_gswm("lthero", KEY, NONCE, 0x5EED, 0, latent, n, 256)
        assert np.array_equal(z >= 0, ref >= 0) and rel_err(z, ref).max() <= REL_TOL
        assert 5.29 < abs(ref[elem]) <= 8.21 and abs(z[elem] - ref[elem]) <= REL_TOL * abs(ref[elem])
        # the same latent inside a large batch (persistent grid, another CTA / lane mapping) is bit-identical
        lo = max(0, latent - 150)
        big = gswm.embed_batch(300, (4, 64, 64), km, 0x5EED, 0, lo, cuda_device)[latent - lo].cpu().numpy().reshape(-1)
        assert np.array_equal(big, z)
    # the refinement formula itself over its whole input range: |z| = -ndtri((m2 + 1/2) 2^-52), m2 = w >> 4
    from scipy.special import ndtri
    rs = np.random.RandomState(9)
    w = np.concatenate([np.array([0, 15, 16, 0xFFFFFFFF, 0xFFFFFFF0, 0x80000000], dtype=np.uint32),
                        rs.randint(0, 1 << 32, size=1 << 20, dtype=np.uint64).astype(np.uint32),
                        (rs.randint(0, 1 << 16, size=1 << 16).astype(np.uint32) << np.uint32(4))])
    d_w = torch.from_numpy(w.view(np.int32)).to(cuda_device)
    d_o = torch.empty(w.size, dtype=torch.float32, device=cuda_device)
    assert gswm._lib.lib().gswm_debug_top_cell(d_w.data_ptr(), w.size, d_o.data_ptr(), None) == 0
    want = -ndtri(((w >> np.uint32(4)).astype(np.float64) + 0.5) * 2.0 ** -52)
    got = d_o.cpu().numpy()
    assert rel_err(got, want).max() <= 5e-7
    assert got.max() <= 8.2096 and got.min() >= 5.2947          # [norm.ppf(1 - 2^-24), norm.ppf(1 - 2^-53)]


def test_philox4x32_known_answers(gswm, cuda_device):
    """The generator core against PUBLISHED vectors: Random123's kat_vectors for philox4x32-10 (the same kernel code runs
    7 rounds in the product), plus the 7-round form against the oracle's restatement on random inputs."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    lib = gswm._lib.lib()
    inp = np.array([list(c) + list(k) for c, k, _ in kat], dtype=np.uint32)
    d_in = torch.from_numpy(inp.view(np.int32)).to(cuda_device)
    d_out = torch.empty((len(kat), 4), dtype=torch.int32, device=cuda_device)
    assert lib.gswm_debug_philox4x32(d_in.data_ptr(), len(kat), 10, d_out.data_ptr(), None) == 0
    got = d_out.cpu().numpy().view(np.uint32)
    for i, (_, _, want) in enumerate(kat):
        assert tuple(int(x) for x in got[i]) == want
    rs = np.random.RandomState(123)
    rnd = rs.randint(0, 1 << 32, size=(4096, 6), dtype=np.uint64).astype(np.uint32)
    d_in = torch.from_numpy(rnd.view(np.int32)).to(cuda_device)
    d_out = torch.empty((4096, 4), dtype=torch.int32, device=cuda_device)
    for rounds in (7, 10):
        assert lib.gswm_debug_philox4x32(d_in.data_ptr(), 4096, rounds, d_out.data_ptr(), None) == 0
        got = d_out.cpu().numpy().view(np.uint32)
        for i in range(0, 4096, 257):
            want = O.philox4x32(rnd[i:i + 1, :4], (int(rnd[i, 4]), int(rnd[i, 5])), rounds)[0]
            assert np.array_equal(got[i], want), (rounds, i)
    assert lib.gswm_philox_rounds() == 7


def test_mt19937_stream_is_numpys(gswm, cuda_device):
    """The device MT19937 against numpy itself (the reference's generator, nodes.py:52-53): RandomState(seed).uniform,
    bit for bit -- > 10^6 consecutive draws (1600+ state regenerations), every position around the 624-word block
    boundaries, extreme seeds, many streams in one launch."""
    u = gswm.mt19937_uniform(42, 1_100_003, device=cuda_device).cpu().numpy()[0]
    want = np.random.RandomState(42).uniform(0, 1, size=u.size)
    assert u.dtype == np.float64 and np.array_equal(u, want)
    assert np.array_equal(u[300:330], want[300:330]) and np.array_equal(u[620:640], want[620:640])   # 312-double regeneration edges
    for seed in (0, 1, 58, 2 ** 31, 2 ** 32 - 1):
        for n in (1, 2, 311, 312, 313, 623, 624, 625, 1000):
            got = gswm.mt19937_uniform(seed, n, device=cuda_device).cpu().numpy()[0]
            assert np.array_equal(got, np.random.RandomState(seed).uniform(size=n)), (seed, n)
    seeds = np.random.RandomState(3).randint(0, 2 ** 32, size=200, dtype=np.uint64)
    got = gswm.mt19937_uniform(seeds, 5000, device=cuda_device).cpu().numpy()
    for i in (0, 1, 57, 199):
        assert np.array_equal(got[i], np.random.RandomState(int(seeds[i])).uniform(size=5000))
    got = gswm.mt19937_uniform(1000, 700, n_streams=3, device=cuda_device).cpu().numpy()       # one int: stream s uses seed + s
    for s_ in range(3):
        assert np.array_equal(got[s_], np.random.RandomState(1000 + s_).uniform(size=700))
    with pytest.raises(ValueError):
        gswm.mt19937_uniform(2 ** 32, 10, device=cuda_device)                                   # numpy raises ValueError too


def test_embed_mt19937_is_the_seeded_reference(gswm, cuda_device, golden, golden_arrays):
    """gswm_embed_mt19937: the reference's seeded embed with nothing uploaded.  Against the reference's own seeded runs
    (ComfyUI node, all shapes incl. multi-tile and ragged ones) and against the injected-uniform path fed numpy's stream."""
    for c in golden["embed_comfy"]:
        n = 4 * (c["width"] // 8) * (c["height"] // 8)
        L = c["message_length"] if c["message_length"] != -1 else gswm.choose_watermark_length(n)
        km = gswm.KeyMaterial.make(KEY, NONCE, gswm.pad_message(c["message"], L // 8), L)
        shape = (4, c["height"] // 8, c["width"] // 8)
        z = gswm.embed_batch_mt19937(c["seed"], 1, shape, km, torch.float32, cuda_device).cpu().numpy().reshape(-1)
        assert sha(np.packbits((z >= 0).astype(np.uint8))) == c["sha256_signs"], c["name"]
        assert rel_err(z[:256], golden_arrays[c["name"] + "_z32_head"]).max() <= REL_TOL, c["name"]
        u = torch.from_numpy(np.random.RandomState(c["seed"]).uniform(size=n)).to(cuda_device)
        zi = gswm.embed_batch_injected(u, shape, km, 1, torch.float32).cpu().numpy().reshape(-1)
        assert np.array_equal(z, zi), c["name"]                  # same uniforms, same arithmetic: bit-identical
        # the Python wrapper takes the generator + injected route for small batches; the fused C entry point itself:
        from gswm.codec import _DeviceJob
        dj = _DeviceJob(km, 1, n, cuda_device)
        zf = torch.empty((1, n), dtype=torch.float32, device=cuda_device)
        assert gswm._lib.lib().gswm_embed_mt19937(C.byref(dj.job), None, c["seed"], zf.data_ptr(), 0,
                                                 torch.cuda.current_stream(cuda_device).cuda_stream) == 0
        assert np.array_equal(zf.cpu().numpy().reshape(-1), z), c["name"]
    # a batch with one seed per latent and per-latent keys, float64 out
    rs = np.random.RandomState(12)
    b, shape, n, L = 9, (4, 64, 64), 16384, 256
    seeds = rs.randint(0, 2 ** 32, size=b, dtype=np.uint64)
    keys, nonces, msgs = rs.bytes(32 * b), rs.bytes(16 * b), rs.bytes(32 * b)
    km = gswm.KeyMaterial.make(keys, nonces, msgs, L)
    z = gswm.embed_batch_mt19937(seeds, b, shape, km, torch.float64, cuda_device).cpu().numpy().reshape(b, n)
    dj = _DeviceJob(km, b, n, cuda_device)
    d_seeds = torch.from_numpy(seeds.astype(np.uint32).view(np.int32)).to(cuda_device)
    zf = torch.empty((b, n), dtype=torch.float64, device=cuda_device)
    assert gswm._lib.lib().gswm_embed_mt19937(C.byref(dj.job), d_seeds.data_ptr(), 0, zf.data_ptr(), 3,
                                             torch.cuda.current_stream(cuda_device).cuda_stream) == 0
    assert np.array_equal(zf.cpu().numpy(), z)                     # fused kernel == generator + injected route
    for i in range(b):
        ref = O.embed(msgs[32 * i:32 * i + 32], keys[32 * i:32 * i + 32], nonces[16 * i:16 * i + 16],
                      np.random.RandomState(int(seeds[i])).uniform(size=n), L)
        assert np.array_equal(z[i] >= 0, ref >= 0) and rel_err(z[i], ref).max() <= 1e-9


# ------------------------------------------------------------------------------------ reference error conditions in K3
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16, torch.float64])
def test_extract_flags_inputs_the_reference_rejects(gswm, cuda_device, golden, dtype):
    """extract.py:83 raises on a NaN (int(nan)), extract.py:86 on any element >= 8.292361075813597 (digit '2').  The extract
    kernel finds both in its one pass: per-latent flags, two counters; everything else about those latents is unchanged."""
    lib = gswm._lib
    km = gswm.KeyMaterial.make(KEY, NONCE, gswm.pad_message("lthero", 32), 256)
    b, n = 40, 16384
    z = gswm.embed_batch(b, (4, 64, 64), km, 5, 0, 0, cuda_device).reshape(b, n).to(dtype)
    clean = gswm.extract_batch(z.clone(), km, want_counts=True)
    assert not clean.flags.any() and clean.counters.cpu().tolist()[4:] == [0, 0]
    big = {torch.float32: 8.29236125946045, torch.float16: 8.296875, torch.bfloat16: 8.3125, torch.float64: 8.292361075813597}[dtype]
    ok = {torch.float32: 8.292360305786133, torch.float16: 8.2890625, torch.bfloat16: 8.25, torch.float64: 8.292361075813595}[dtype]
    from scipy.stats import norm
    assert int(norm.cdf(big) * 2) == 2 and int(norm.cdf(ok) * 2) == 1            # the type's first rejected / last accepted value
    want = np.zeros(b, dtype=np.uint8)
    plan = {1: (0, float("nan")), 2: (n - 1, float("nan")), 3: (7777, -float("nan")), 5: (0, big), 6: (n - 1, float("inf")),
            7: (12345, 9.0), 8: (100, ok), 9: (200, -float("inf")), 10: (300, -9.0), 11: (5, float("nan")), 12: (8191, big),
            13: (8192, big), 39: (16383, float("nan"))}
    for row, (col, val) in plan.items():
        z[row, col] = val
        want[row] = lib.FLAG_NAN if val != val else (lib.FLAG_RANGE if val >= big else 0)
    z[11, 9] = big                                                             # NaN and an oversized value: NaN wins (it raises first)
    res = gswm.extract_batch(z, km, want_counts=True)
    assert np.array_equal(res.flags.cpu().numpy(), want)
    c = res.counters.cpu().tolist()
    assert c[4] == int((want == lib.FLAG_NAN).sum()) and c[5] == int((want == lib.FLAG_RANGE).sum()) and c[3] == b
    # the oracle (= the reference) raises exactly for the flagged rows, and agrees on the counts of all the others
    zh = z.double().cpu().numpy() if dtype == torch.float64 else z.float().cpu().numpy()
    for i in range(b):
        if want[i]:
            with pytest.raises(ValueError):
                O.vote_counts(zh[i], KEY, NONCE, 256)
        else:
            assert np.array_equal(res.counts[i].cpu().numpy().astype(np.uint32), O.vote_counts(zh[i], KEY, NONCE, 256))
    # the golden cases through the two front-ends
    import io
    from gswm import extract as gx
    args = types.SimpleNamespace(key=KEY, nonce=NONCE, l=1, message_length=256, original_message_hex=(b"lthero" + bytes(26)).hex())
    for bad in golden["quantise_raises"]:
        t = z[0].clone().reshape(1, 4, 64, 64)
        t[0, 0, 0, 0] = float(bad["z"])
        if dtype != torch.float64 and float(bad["z"]) == 8.292361075813597:
            continue                                                            # not representable in the narrower types
        with pytest.raises(ValueError):
            gx.recover_exactracted_message(t.cpu(), args)
    buf = io.StringIO()
    strings, accs, avg = gx.evaluate_latents([f"img{i}.png" for i in range(b)], z.cpu(), args, buf)
    text = buf.getvalue()
    for i in range(b):
        if want[i]:
            assert strings[i] is None and accs[i] is None and f"Error processing img{i}.png: " in text
        else:
            assert accs[i] == 1.0 and f"img{i}.png, Bit Accuracy, 1.0" in text
    assert avg == 1.0 and "Average Bit Accuracy, 1.0" in text


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("shape,L", [((4, 8, 16), 8), ((4, 8, 16), 4), ((4, 8, 16), 1), ((4, 8, 16), 2), ((4, 8, 16), 16),
                                     ((4, 6, 8), 12), ((4, 6, 8), 24), ((4, 10, 10), 100), ((4, 64, 64), 8), ((4, 64, 64), 20480 // 5),
                                     ((4, 96, 64), 24), ((4, 64, 64), 16), ((4, 64, 64), 16384 // 4), ((4, 72, 72), 81)])
def test_extract_any_message_length(gswm, cuda_device, dtype, shape, L):
    """extract.py:195 `--message_length` is an arbitrary integer: any length that divides the latent decodes (segments of
    L bits, extract.py:91-98), not only multiples of 32.  Counts, messages (rows of ceil(L/8) bytes, unused low bits zero)
    and matched bits against the oracle."""
    n = int(np.prod(shape))
    assert n % L == 0
    rs = np.random.RandomState(n * 31 + L)
    b = 5
    zn = torch.from_numpy(rs.standard_normal((b, *shape)).astype(np.float32)).to(dtype)
    ref_msg = rs.bytes((L + 7) // 8)
    km = gswm.KeyMaterial.make(KEY, NONCE, ref_msg, L)
    res = gswm.extract_batch(zn.to(cuda_device), km, want_counts=True)
    zh = zn.float().numpy()
    ref_bits = np.unpackbits(np.frombuffer(ref_msg, np.uint8))[:L]
    total = 0
    for i in range(b):
        assert np.array_equal(res.counts[i].cpu().numpy().astype(np.uint32), O.vote_counts(zh[i], KEY, NONCE, L)), i
        bits = O.recover_message_bits(zh[i], KEY, NONCE, L)
        assert O.bits_to_bytes(bits) == res.messages[i].cpu().numpy().tobytes()
        assert res.bit_strings()[i] == O.recover_message(zh[i], KEY, NONCE, L)
        m = int((bits == ref_bits).sum())
        assert int(res.matched[i]) == m
        total += m
    assert res.counters.cpu().tolist() == [total, b * L, int((res.matched == L).sum()), b, 0, 0]
    # the reference-named function with the same length
    from gswm import extract as gx
    args = types.SimpleNamespace(key=KEY, nonce=NONCE, l=1, message_length=L)
    assert gx.recover_exactracted_message(zn[:1], args) == O.recover_message(zh[0], KEY, NONCE, L)
    with pytest.raises(IndexError):
        gx.recover_exactracted_message(zn[:1], types.SimpleNamespace(key=KEY, nonce=NONCE, l=1, message_length=n - 1))


def test_keys_in_flight_flag_changes_nothing_but_the_ordering(gswm, cuda_device):
    """GSWM_JOB_KEYS_IN_FLIGHT moves every key-material read behind the grid dependency wait (for callers whose key
    material is produced by a programmatic-launch kernel just ahead in the stream).  Same results, both kernels, both
    key modes -- with the key material written by a device kernel (torch copy) immediately before each launch."""
    import ctypes as C
    lib = gswm._lib.lib()
    rs = np.random.RandomState(6)
    for per in (0, 1):
        b, n, L = 300, 16384, 256
        rows = b if per else 1
        src = torch.from_numpy(np.frombuffer(rs.bytes(80 * rows), np.uint8).copy()).to(cuda_device)
        flat = torch.zeros_like(src)
        outs = []
        for flag in (0, gswm._lib.JOB_KEYS_IN_FLIGHT):
            z = torch.empty((b, n), dtype=torch.float32, device=cuda_device)
            msgs = torch.empty((b, 32), dtype=torch.uint8, device=cuda_device)
            ctr = torch.zeros(6, dtype=torch.int64, device=cuda_device)
            job = gswm._lib.Job(b, n, L, per | flag, flat.data_ptr(), flat.data_ptr() + 32 * rows, flat.data_ptr() + 48 * rows)
            sp = torch.cuda.current_stream().cuda_stream
            for _ in range(3):                                   # back to back: embed, extract, embed, extract, ...
                flat.zero_()
                flat.copy_(src)                                  # key material lands just ahead of the launches
                assert lib.gswm_embed(C.byref(job), 9, 0, 0, z.data_ptr(), sp) == 0
                assert lib.gswm_extract(C.byref(job), z.data_ptr(), 0, msgs.data_ptr(), None, None, None, ctr.data_ptr(), sp) == 0
            outs.append((z.clone(), msgs.clone(), ctr.clone()))
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])
        assert outs[0][2].tolist() == [3 * b * L, 3 * b * L, 3 * b, 3 * b, 0, 0]
        want = src[48 * rows:].cpu().numpy().tobytes()
        assert outs[0][1].cpu().numpy().tobytes() == (want if per else want * b)


# ------------------------------------------------------------------------------------ multi-GPU exchange (one GPU is enough to run it)
def test_comm_allreduce_two_ranks_on_one_device(gswm, cuda_device):
    """gswm_comm with both ranks in this process (gswm_comm_connect_local) and on this one GPU: the mailbox exchange is
    the same code that runs over NVLink -- stores into the peer's mailbox, release flag, acquire spin, sum.  The two
    ranks' kernels run concurrently on two streams.  Stand-alone all-reduce, then the form fused into the extract
    kernel, several epochs (the mailboxes are double-buffered by epoch parity)."""
    import ctypes as C
    lib = gswm._lib.lib()
    dev = cuda_device.index or 0
    hs = [C.c_void_p(), C.c_void_p()]
    for r in range(2):
        assert lib.gswm_comm_create(C.byref(hs[r]), dev, r, 2, None) == 0
    arr = (C.c_void_p * 2)(hs[0], hs[1])
    assert lib.gswm_comm_connect_local(arr, 2) == 0
    streams = [torch.cuda.Stream(cuda_device), torch.cuda.Stream(cuda_device)]
    bufs = [torch.zeros(6, dtype=torch.int64, device=cuda_device) for _ in range(2)]
    for epoch in range(5):
        vals = [[10 * epoch + r + i for i in range(6)] for r in range(2)]
        for r in range(2):
            bufs[r].copy_(torch.tensor(vals[r], dtype=torch.int64))
        torch.cuda.synchronize()
        for r in range(2):
            assert lib.gswm_comm_allreduce_counters(hs[r], bufs[r].data_ptr(), 6, streams[r].cuda_stream) == 0
        torch.cuda.synchronize()
        want = [vals[0][i] + vals[1][i] for i in range(6)]
        assert bufs[0].tolist() == want and bufs[1].tolist() == want
        assert lib.gswm_comm_status(hs[0]) == 0 and lib.gswm_comm_status(hs[1]) == 0
    # fused into K3: rank r decodes its own shard; each rank's `reduced` holds the sum of both ranks' ACCUMULATED counters
    msg = gswm.pad_message("lthero", 32)
    km = gswm.KeyMaterial.make(KEY, NONCE, msg, 256)
    from gswm.codec import _DeviceJob
    shards = [(0, 700), (700, 1000)]
    zs = [gswm.embed_batch(hi - lo, (4, 64, 64), km, 3, 0, lo, cuda_device) for lo, hi in shards]
    zs[1][5, 0, 0, 0] = float("nan")                                # one rejected latent on rank 1
    ctrs = [torch.zeros(6, dtype=torch.int64, device=cuda_device) for _ in range(2)]
    reds = [torch.full((6,), -1, dtype=torch.int64, device=cuda_device) for _ in range(2)]
    outs = [torch.empty((hi - lo, 32), dtype=torch.uint8, device=cuda_device) for lo, hi in shards]
    jobs = [_DeviceJob(km, hi - lo, 16384, cuda_device) for lo, hi in shards]
    torch.cuda.synchronize()
    for step in range(3):
        for r in range(2):
            if step < 2:                                          # plain launches accumulate ...
                rc = lib.gswm_extract(C.byref(jobs[r].job), zs[r].data_ptr(), 0, outs[r].data_ptr(), None, None, None,
                                      ctrs[r].data_ptr(), streams[r].cuda_stream)
            else:                                                 # ... the last one carries the exchange
                rc = lib.gswm_extract_allreduce(C.byref(jobs[r].job), zs[r].data_ptr(), 0, outs[r].data_ptr(), None, None, None,
                                                ctrs[r].data_ptr(), hs[r], reds[r].data_ptr(), streams[r].cuda_stream)
            assert rc == 0
    torch.cuda.synchronize()
    assert ctrs[0].tolist() == [3 * 700 * 256, 3 * 700 * 256, 3 * 700, 3 * 700, 0, 0]
    # the latent holding a NaN is flagged (counter 4) AND still decoded (NaN -> bit 0, one vote of 64): gswm.h, d_flags
    assert ctrs[1].tolist() == [3 * 300 * 256, 3 * 300 * 256, 3 * 300, 3 * 300, 3, 0]
    want = (ctrs[0] + ctrs[1]).tolist()
    assert reds[0].tolist() == want and reds[1].tolist() == want
    assert lib.gswm_comm_status(hs[0]) == 0
    # a rank with an empty shard decodes nothing and still takes part: its accumulated counters are summed with the peer's
    empty = gswm._lib.Job(0, 16384, 256, 0, jobs[1].job.keys, jobs[1].job.nonces, jobs[1].job.msgs)
    assert lib.gswm_extract_allreduce(C.byref(jobs[0].job), zs[0].data_ptr(), 0, outs[0].data_ptr(), None, None, None,
                                      ctrs[0].data_ptr(), hs[0], reds[0].data_ptr(), streams[0].cuda_stream) == 0
    assert lib.gswm_extract_allreduce(C.byref(empty), None, 0, outs[1].data_ptr(), None, None, None,
                                      ctrs[1].data_ptr(), hs[1], reds[1].data_ptr(), streams[1].cuda_stream) == 0
    torch.cuda.synchronize()
    assert ctrs[0].tolist()[:4] == [4 * 700 * 256, 4 * 700 * 256, 4 * 700, 4 * 700] and ctrs[1].tolist()[3] == 3 * 300
    want = (ctrs[0] + ctrs[1]).tolist()
    assert reds[0].tolist() == want and reds[1].tolist() == want
    # argument errors do not consume an epoch (the ranks would fall out of step)
    bad = gswm._lib.Job(1, 16384, 640, 0, 16, 16, None)
    assert lib.gswm_extract_allreduce(C.byref(bad), zs[0].data_ptr(), 0, outs[0].data_ptr(), None, None, None, ctrs[0].data_ptr(),
                                      hs[0], reds[0].data_ptr(), None) == -3
    for r in range(2):
        assert lib.gswm_comm_allreduce_counters(hs[r], bufs[r].data_ptr(), 6, streams[r].cuda_stream) == 0
    torch.cuda.synchronize()
    assert bufs[0].tolist() == bufs[1].tolist() and lib.gswm_comm_status(hs[1]) == 0      # still in step
    for h in hs:
        lib.gswm_comm_destroy(h)
    # a one-rank communicator is its own peer: the Python wrapper outside torch.distributed
    comm = gswm.Comm(cuda_device)
    t = torch.arange(6, dtype=torch.int64, device=cuda_device)
    assert comm.allreduce_counters(t).tolist() == [0, 1, 2, 3, 4, 5] and comm.status() == 0
    res = gswm.extract_batch(zs[0], km, comm=comm)
    assert res.reduced.tolist() == res.counters.tolist() == [700 * 256, 700 * 256, 700, 700, 0, 0]
    comm.close()
