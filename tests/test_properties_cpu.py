"""Property tests (hypothesis) of the host logic and of the oracle, CPU only: things that must hold for EVERY input, not just the
golden cases -- framing against the oracle, sharding as a partition, bit accuracy against the reference's restatement, the
vote's erasure tolerance, and injectivity of the Philox counter layout."""
import os
import sys

import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from oracle import gs_oracle as O  # noqa: E402

KEY = bytes.fromhex(O.DEFAULT_KEY_HEX)
NONCE = bytes.fromhex(O.DEFAULT_NONCE_HEX)
FAST = settings(max_examples=60, deadline=None)


@pytest.fixture(scope="module")
def gswm():
    import gswm as g
    return g


@FAST
@given(msg=st.text(min_size=1, max_size=80), n_bytes=st.sampled_from([4, 8, 16, 32, 64, 128]), repeat=st.booleans())
def test_pad_message_is_the_oracles_for_any_text(gswm, msg, n_bytes, repeat):
    """gs_insert.py:9-20 / v1.5.2:29-47: UTF-8 bytes, zero padded or cut -- also in the middle of a code point."""
    if repeat and n_bytes % 4:
        return
    want = O.frame_message(msg, 8 * n_bytes, 8 * n_bytes, use_repeat=repeat)[0] if repeat else O.pad_message(msg, n_bytes)
    got = gswm.pad_message(msg, n_bytes, use_repeat=repeat)
    assert got == want and len(got) == n_bytes


@FAST
@given(n=st.integers(0, 10 ** 7), world=st.integers(1, 64))
def test_shard_range_is_a_balanced_partition(gswm, n, world):
    edges = [gswm.sharding.shard_range(n, r, world) for r in range(world)]
    assert edges[0][0] == 0 and edges[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))                       # contiguous, no gap, no overlap
    sizes = [hi - lo for lo, hi in edges]
    assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)      # balanced, the larger shards first


@FAST
@given(hexmsg=st.text(alphabet="0123456789abcdef", min_size=1, max_size=70), bits=st.text(alphabet="01", min_size=1, max_size=300))
def test_calculate_bit_accuracy_is_the_references_for_any_strings(gswm, hexmsg, bits):
    """extract.py:103-110, including its silent truncation to the shorter string."""
    from gswm import extract
    got, want = extract.calculate_bit_accuracy(hexmsg, bits), O.calculate_bit_accuracy(hexmsg, bits)
    assert got[0] == want[0] and got[1] == want[1]


@settings(max_examples=25, deadline=None)
@given(data=st.data(), l_bits=st.sampled_from([32, 64, 128, 256]), copies=st.integers(3, 9))
def test_majority_vote_survives_any_minority_of_flipped_copies(data, l_bits, copies):
    """extract.py:91-99: a position decodes correctly as long as strictly more than half of its copies agree; flipping the signs
    of fewer than half of the copies of every position (any pattern) leaves the message intact, flipping exactly half of an even
    number of copies turns a 1 into a 0 (tie -> 0)."""
    n = l_bits * copies
    msg = data.draw(st.binary(min_size=l_bits // 8, max_size=l_bits // 8))
    u = np.random.RandomState(data.draw(st.integers(0, 2 ** 31))).uniform(size=n)
    z = O.embed(msg, KEY, NONCE, u, l_bits)
    want = "".join(format(b, "08b") for b in msg)
    assert O.recover_message(z, KEY, NONCE, l_bits) == want
    k = (copies - 1) // 2                                                             # a strict minority of the copies
    flips = np.zeros((copies, l_bits), dtype=bool)
    for p in range(l_bits):
        flips[data.draw(st.permutations(range(copies)))[:k], p] = True
    zf = np.where(flips.reshape(-1), -z, z)
    zf[zf == 0] = -1e-3                                                               # (-0.0 would still quantise to 1)
    assert O.recover_message(zf, KEY, NONCE, l_bits) == want


@FAST
@given(offset=st.integers(0, 2 ** 62 - 1), latent=st.integers(0, 2 ** 40), tiles=st.integers(1, 16), seed=st.integers(0, 2 ** 64 - 1))
def test_v4_counters_are_distinct_for_distinct_positions(offset, latent, tiles, seed):
    """Uniforms v4: (latent, tile, super-iteration, lane, call) -> Philox counter is injective (T < 2^54 splits over two words
    next to the 8-bit lane index and the 2-bit call index), so no two float4 groups ever share a random word."""
    rs = np.random.RandomState(seed % (2 ** 32))
    k = 512
    tile = rs.randint(0, tiles, size=k).astype(np.uint64)
    s_ = rs.randint(0, 4, size=k).astype(np.uint64)
    tid = rs.randint(0, 256, size=k).astype(np.uint64)
    seen = {}
    for call in range(4):
        ctr = O._gswm_counters(offset, latent, tiles, tile, s_, tid, call)
        for i in range(k):
            key = tuple(int(x) for x in ctr[i])
            pos = (int(tile[i]), int(s_[i]), int(tid[i]), call)
            assert seen.setdefault(key, pos) == pos
    # neighbouring latents do not collide with this one either
    other = O._gswm_counters(offset, latent + 1, tiles, tile, s_, tid, 0)
    assert not ({tuple(int(x) for x in r) for r in other} & set(seen))
