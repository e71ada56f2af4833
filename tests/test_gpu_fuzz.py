"""Randomised parity (hypothesis) of the CUDA path against the oracle: latent shapes the ComfyUI node can emit (any multiple of
8 pixels: ragged tiles, several tiles), message lengths, batch sizes, first-latent offsets, seeds, shared or per-latent keys,
every input type of the extract kernel.  Embed: bucket membership exact, values within 1e-6 relative; extract: per-position
counts, decoded bytes, matched bits and counters bit-exact."""
import os

import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import gs_oracle as O

pytestmark = pytest.mark.gpu
REL_TOL = 1e-6


@pytest.fixture(scope="module")
def gswm(cuda_device):
    import gswm as g
    g.build()
    g._lib.lib()   # raises if the extension cannot be loaded: there is no fallback path
    return g


def _rel(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    return np.abs(got - ref) / np.abs(ref)


@settings(max_examples=int(os.environ.get("GSWM_FUZZ_EXAMPLES", "30")), deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(h=st.integers(1, 24), w=st.integers(1, 24), lsel=st.integers(0, 10 ** 6), b=st.integers(1, 40),
       first=st.integers(0, 2 ** 40), seed=st.integers(0, 2 ** 64 - 1), offset=st.integers(0, 2 ** 62 - 1),
       per_latent=st.booleans(), dtype=st.sampled_from(["f32", "f16", "bf16", "f64"]), sigma=st.sampled_from([0.0, 0.5, 3.0]),
       rs_seed=st.integers(0, 2 ** 31))
def test_random_jobs_match_the_oracle(gswm, cuda_device, h, w, lsel, b, first, seed, offset, per_latent, dtype, sigma, rs_seed):
    shape = (4, 8 * h, 8 * w)                                   # latent of an (64 h) x (64 w) pixel image: n = 256 h w elements
    n = int(np.prod(shape))
    lengths = [L for L in (32, 64, 96, 128, 160, 256, 320, 512, 1024, 2048) if L <= n and n % L == 0]
    L = lengths[lsel % len(lengths)]
    rs = np.random.RandomState(rs_seed)
    rows = b if per_latent else 1
    keys, nonces, msgs = rs.bytes(32 * rows), rs.bytes(16 * rows), rs.bytes((L // 8) * rows)
    km = gswm.KeyMaterial.make(keys, nonces, msgs, L)
    z = gswm.embed_batch(b, shape, km, seed, offset, first, cuda_device)
    zh = z.cpu().numpy().reshape(b, n)
    row = lambda i: (msgs[(L // 8) * i:(L // 8) * (i + 1)], keys[32 * i:32 * i + 32], nonces[16 * i:16 * i + 16]) if per_latent \
        else (msgs, keys, nonces)
    for i in sorted({0, b - 1, int(rs.randint(0, b))}):
        m, k, no = row(i)
        ref = O.embed_gswm(m, k, no, seed, offset, first + i, n, L)
        assert np.array_equal(zh[i] >= 0, ref >= 0), (shape, L, i)
        assert _rel(zh[i], ref).max() <= REL_TOL, (shape, L, i)
    tdt = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16, "f64": torch.float64}[dtype]
    if (n * torch.empty((), dtype=tdt).element_size()) % 16:
        tdt = torch.float32                                      # rows must be 16-byte multiples (the drop-ins upcast such inputs too)
    noisy = (z + sigma * torch.randn(z.shape, device=cuda_device, generator=torch.Generator(cuda_device).manual_seed(rs_seed)))
    noisy = noisy.clamp(max=8.0).to(tdt)
    res = gswm.extract_batch(noisy, km, want_counts=True)
    nh = noisy.double().cpu().numpy().reshape(b, n)              # exact upcast: the oracle sees the values the kernel sees
    total_matched, exact = 0, 0
    for i in range(b):
        m, k, no = row(i)
        want = O.vote_counts(nh[i], k, no, L)
        assert np.array_equal(res.counts[i].cpu().numpy().astype(np.uint32), want), (shape, L, dtype, i)
        bits = (want > (n // L) / 2).astype(np.uint8)
        assert np.packbits(bits).tobytes() == res.messages[i].cpu().numpy().tobytes()
        matched = int((bits == np.unpackbits(np.frombuffer(m, np.uint8))).sum())
        assert matched == int(res.matched[i])
        total_matched += matched
        exact += matched == L
    assert res.counters.tolist() == [total_matched, b * L, exact, b, 0, 0]
    assert not res.flags.any()
    if sigma == 0.0:
        assert exact == b                                        # round trip
