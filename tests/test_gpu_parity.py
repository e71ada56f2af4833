"""Parity of the CUDA path (libgswm.so through its C ABI) against the oracle and the golden vectors.

Run on the B200 box:  python -m pytest tests -m gpu -x -q
Bars (BASELINE.json north_star): keystream, bucket membership (sign), vote counts, decoded messages and
bit counts BIT-EXACT; latents within 1e-6 relative of scipy's float64 norm.ppf.
"""
import ctypes as C
import hashlib
import types

import numpy as np
import pytest
import torch

from oracle import gs_oracle as O

pytestmark = pytest.mark.gpu

REL_TOL = 1e-6   # north_star: latents within 1e-6 relative of scipy's float64 norm.ppf
KEY = bytes.fromhex(O.DEFAULT_KEY_HEX)
NONCE = bytes.fromhex(O.DEFAULT_NONCE_HEX)


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
    assert list(c) == [matched_total, b * L, int(sum(int(x) == L for x in res.matched.cpu())), b]


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
    assert list(res.counters.cpu().numpy()) == [4096 * 256, 4096 * 256, 4096, 4096]
    # moments of the watermarked noise: it must still look like N(0, 1)
    assert abs(float(z.mean())) < 2e-3 and abs(float(z.std()) - 1.0) < 2e-3
    # sigma = 0.325 regime: ~90 % of signs agree, every message still decodes
    noisy = z + 0.325 * torch.randn(z.shape, device=cuda_device, generator=torch.Generator(cuda_device).manual_seed(99))
    agree = float(((noisy >= 0) == (z >= 0)).float().mean())
    assert 0.89 < agree < 0.91
    res = gswm.extract_batch(noisy, km, want_counts=True)
    assert res.bit_accuracy() == 1.0
    sub = noisy[:16].cpu().numpy()
    for i in range(16):
        assert np.array_equal(res.counts[i].cpu().numpy().astype(np.uint32), O.vote_counts(sub[i], KEY, NONCE, 256))
    # sigma = 4.3 regime: ~90 % of the DECODED bits survive; counts / messages / matched bits of a
    # subset bit-exact against the oracle, and the batch counters equal the sum of the per-latent figures
    heavy = (z + 4.3 * torch.randn(z.shape, device=cuda_device, generator=torch.Generator(cuda_device).manual_seed(7))).clamp(max=8.0)
    res = gswm.extract_batch(heavy, km, want_counts=True)
    assert 0.88 < res.bit_accuracy() < 0.92
    c = res.counters.cpu().numpy()
    assert int(c[0]) == int(res.matched.sum()) and int(c[1]) == 4096 * 256 and int(c[3]) == 4096
    assert int(c[2]) == int((res.matched == 256).sum())
    sub = heavy[:16].cpu().numpy()
    ref_bits = np.unpackbits(np.frombuffer(msg, np.uint8))
    for i in range(16):
        assert np.array_equal(res.counts[i].cpu().numpy().astype(np.uint32), O.vote_counts(sub[i], KEY, NONCE, 256))
        bits = O.recover_message_bits(sub[i], KEY, NONCE, 256)
        assert O.bits_to_bytes(bits) == res.messages[i].cpu().numpy().tobytes()
        assert int(res.matched[i]) == int((bits == ref_bits).sum())
    del z, noisy, heavy
    zx = gswm.embed_batch(512, (4, 128, 128), km, 1, 0, 0, cuda_device)
    rx = gswm.extract_batch(zx, km)
    assert list(rx.counters.cpu().numpy()) == [512 * 256, 512 * 256, 512, 512]


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
        assert list(res.counters.cpu().numpy()) == [b * L, b * L, b, b]
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
    assert list(res.counters.cpu().numpy()) == [B * 256, B * 256, B, B]
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
    sub = tail[:4].cpu().numpy().reshape(4, n)
    for i in range(4):
        ref = O.embed_gswm("lthero", KEY, NONCE, 0x5EED, 0, B - 256 + i, n, 256)
        assert np.array_equal(sub[i] >= 0, ref >= 0) and rel_err(sub[i], ref).max() <= REL_TOL
    # no two latents share their noise: |z| differs between latents (same signs, different magnitudes)
    assert not torch.equal(z[0].abs(), z[1].abs()) and not torch.equal(z[0].abs(), z[B - 1].abs())


def test_config5_per_latent_keys_1m(gswm, cuda_device):
    """BASELINE config 5 at full size: 1 M SD-2.1 latents, each with its own key / nonce / message (RandomState(2025)),
    streamed through the device API in 8 chunks of 131 072 (8.6 GB each).  Every decoded message must equal the
    message embedded in that latent; a 64-latent subset is checked against the oracle through `cryptography`."""
    free, _ = torch.cuda.mem_get_info(cuda_device)
    total, chunk, shape, n, L = 1 << 20, 1 << 17, (4, 64, 64), 16384, 256
    if free < chunk * n * 4 + (4 << 30):
        pytest.skip("needs ~13 GB of free HBM")
    rs = np.random.RandomState(2025)
    keys = np.frombuffer(rs.bytes(32 * total), np.uint8).reshape(total, 32)
    nonces = np.frombuffer(rs.bytes(16 * total), np.uint8).reshape(total, 16)
    msgs = np.frombuffer(rs.bytes(32 * total), np.uint8).reshape(total, 32)
    out = torch.empty((chunk, *shape), dtype=torch.float32, device=cuda_device)
    counters = torch.zeros(4, dtype=torch.int64, device=cuda_device)
    for c in range(total // chunk):
        sl = slice(c * chunk, (c + 1) * chunk)
        km = gswm.KeyMaterial.make(keys[sl], nonces[sl], msgs[sl], L)
        gswm.embed_batch(chunk, shape, km, 2025, 0, c * chunk, cuda_device, out=out)
        res = gswm.extract_batch(out, km, counters=counters)
        assert torch.equal(res.messages.cpu(), torch.from_numpy(msgs[sl].copy()))
        if c in (0, 7):
            idx = [0, 1, chunk // 2, chunk - 1]
            zs = out[idx].cpu().numpy().reshape(len(idx), n)
            for j, i in enumerate(idx):
                g = c * chunk + i
                k, no, m = keys[g].tobytes(), nonces[g].tobytes(), msgs[g].tobytes()
                y = O.bucket_bits(O.frame_message(m, n, L)[1], k, no, keystream=O.chacha20_keystream_lib)[:n]
                assert np.array_equal(zs[j] >= 0, y == 1)
                ref = O.embed_gswm(m, k, no, 2025, 0, g, n, L)
                assert rel_err(zs[j], ref).max() <= REL_TOL
    assert list(counters.cpu().numpy()) == [total * L, total * L, total, total]


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
    msgs, cnt, matched, counters = pipe.extract(noisy, km, want_counts=True)
    r = gswm.extract_batch(noisy.to(cuda_device), km, want_counts=True)
    assert np.array_equal(msgs, r.messages.cpu().numpy())
    assert np.array_equal(cnt, r.counts.cpu().numpy())
    assert np.array_equal(matched, r.matched.cpu().numpy())
    assert np.array_equal(counters, r.counters.cpu().numpy())
    # fp16 host input, per-latent keys
    rs = np.random.RandomState(4)
    kmp = gswm.KeyMaterial.make(rs.bytes(32 * 27), rs.bytes(16 * 27), rs.bytes(32 * 27), 256)
    pipe.embed(out, kmp, 6)
    m2, _, mt2, c2 = pipe.extract(out.half(), kmp)
    assert m2.tobytes() == kmp.msgs.tobytes() and list(c2) == [27 * 256, 27 * 256, 27, 27]
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


def test_argument_errors(gswm, cuda_device):
    km = gswm.KeyMaterial.make(KEY, NONCE, bytes(32), 256)
    with pytest.raises(ValueError):
        gswm.embed_batch(1, (3, 5, 5), km, 0, device=cuda_device)          # not a multiple of 4
    with pytest.raises(ValueError):
        gswm.extract_batch(torch.zeros((1, 4, 96, 64), device=cuda_device), gswm.KeyMaterial.make(KEY, NONCE, None, 640))
    lib = gswm._lib.lib()
    job = gswm._lib.Job(1, 16384, 256, 0, 1, 1, 1)
    assert lib.gswm_embed(C.byref(job), 0, 0, 0, 16, 16, None) == -7              # misaligned key pointer
    job = gswm._lib.Job(1, 16384, 250, 0, 16, 16, 16)
    assert lib.gswm_embed(C.byref(job), 0, 0, 0, 16, 16, None) == -3
    job = gswm._lib.Job(1, 1002, 32, 0, 16, 16, 16)
    assert lib.gswm_embed(C.byref(job), 0, 0, 0, 16, 16, None) == -2
    assert lib.gswm_embed(None, 0, 0, 0, 16, 16, None) == -1
    assert "multiple of 4" in gswm._lib.strerror(-2)


def test_empty_batches(gswm, cuda_device):
    """Zero latents is a valid job everywhere: nothing is launched, outputs are empty, counters stay zero."""
    km = gswm.KeyMaterial.make(KEY, NONCE, gswm.pad_message("lthero", 32), 256)
    before = gswm.launch_count()
    z = gswm.embed_batch(0, (4, 64, 64), km, 1, device=cuda_device)
    assert tuple(z.shape) == (0, 4, 64, 64) and z.dtype == torch.float32 and z.is_cuda
    res = gswm.extract_batch(z, km, want_counts=True)
    assert tuple(res.messages.shape) == (0, 32) and tuple(res.counts.shape) == (0, 256) and tuple(res.matched.shape) == (0,)
    assert res.counters.cpu().tolist() == [0, 0, 0, 0] and res.bit_strings() == [] and np.isnan(res.bit_accuracy())
    pipe = gswm.HostPipe(cuda_device.index or 0, max_elems=16384, chunk_latents=4)
    out = pipe.embed(np.empty((0, 4, 64, 64), dtype=np.float32), km, 1)
    msgs, cnt, matched, counters = pipe.extract(out, km, want_counts=True)
    assert msgs.shape == (0, 32) and cnt.shape == (0, 256) and matched.shape == (0,) and counters.tolist() == [0, 0, 0, 0]
    pipe.close()
    assert gswm.launch_count() == before
    # the C ABI itself: n_latents == 0 is success, before any launch
    lib = gswm._lib.lib()
    flat = torch.zeros(96, dtype=torch.uint8, device=cuda_device)
    buf = torch.zeros(64, dtype=torch.float32, device=cuda_device)
    job = gswm._lib.Job(0, 16384, 256, 0, flat.data_ptr(), flat.data_ptr() + 32, flat.data_ptr() + 48)
    assert lib.gswm_embed(C.byref(job), 0, 0, 0, buf.data_ptr(), None, None) == 0
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
