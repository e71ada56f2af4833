"""CPU oracle for the Gaussian-Shading embed/extract hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or the
reported CPU baseline.  The product path (``gswm`` + ``libgswm.so``) never
imports this module and has no CPU fallback.

What it restates (file:line are relative to the reference repository):

* message framing ............ gs_insert.py:9-23, nodes.py:56-87, v1.5.2:29-47
* key / nonce resolution ..... gs_insert.py:27-42, extract.py:200-204
* ChaCha20 (third party) ..... called at gs_insert.py:45-47, extract.py:77-78,87.
  The algorithm lives in PyPI ``cryptography`` (unpinned in requirements.txt:2;
  48.0.0 / OpenSSL here).  Its 16-byte "nonce" is the original DJB layout:
  state words 12..13 are a 64-bit little-endian block counter seeded from
  nonce[0:8], words 14..15 are nonce[8:16].  Restated below in numpy and pinned
  against ``cryptography`` itself and RFC 7539 section 2.3.2.
* bit expansion .............. gs_insert.py:49,58-60 (MSB first inside a byte)
* bucket sample .............. gs_insert.py:62-64: z = norm.ppf((u + y) / 2)
  (``scipy.stats.norm.ppf`` == ``scipy.special.ndtri``, third party, unpinned)
* scatter .................... gs_insert.py:56,65 (C-order flat index)
* sign quantise .............. extract.py:82-84: int(norm.cdf(z) * 2)
* pack + decrypt ............. extract.py:86-89
* majority vote .............. extract.py:91-99 (strict majority, tie -> '0')
* bit accuracy ............... extract.py:103-110

Parity pinning: the reference has no tests or golden vectors of its own
(SURVEY.md section 4), so this oracle is pinned against the reference *itself*,
imported verbatim from /root/reference by ``tests/golden/make_golden.py``; the
resulting vectors are committed under ``tests/golden/`` and checked by
``tests/test_oracle_golden.py``.

Two flavours are provided: vectorised numpy (used for thousands of latents) and
``*_scalar`` pure-Python loops that follow the reference statement by statement
(small cases only).
"""
from __future__ import annotations

import numpy as np
from scipy.special import ndtr, ndtri

DEFAULT_KEY_HEX = "5822ff9cce6772f714192f43863f6bad1bf54b78326973897e6b66c3186b77a7"  # README.md:61
DEFAULT_NONCE_HEX = "05072fd1c2265f6f2e2a4080a2bfbdd8"  # README.md:67

# extract.py:83 -- int(norm.cdf(z) * 2) is 1 for every float64 z >= this value,
# because cdf(z) rounds to exactly 0.5 there (probed).
CDF_HALF_THRESHOLD = -6.957291061679417e-17
# extract.py:83-86 -- for z >= this, int(cdf*2) == 2 and the reference crashes.
CDF_ONE_THRESHOLD = 8.292361075813597


# --------------------------------------------------------------------------
# ChaCha20 (DJB layout, 64-bit counter) -- numpy restatement
# --------------------------------------------------------------------------
_SIGMA = np.frombuffer(b"expand 32-byte k", dtype="<u4")


def _rotl(x, n):
    return (x << np.uint32(n)) | (x >> np.uint32(32 - n))


def _qr(s, a, b, c, d):
    s[a] += s[b]; s[d] ^= s[a]; s[d] = _rotl(s[d], 16)
    s[c] += s[d]; s[b] ^= s[c]; s[b] = _rotl(s[b], 12)
    s[a] += s[b]; s[d] ^= s[a]; s[d] = _rotl(s[d], 8)
    s[c] += s[d]; s[b] ^= s[c]; s[b] = _rotl(s[b], 7)


def chacha20_blocks(key: bytes, nonce16: bytes, nblocks: int, first_block: int = 0) -> np.ndarray:
    """Keystream words, shape (nblocks, 16) uint32, for one (key, nonce).

    Block j uses the 64-bit counter  LE64(nonce16[0:8]) + first_block + j  (mod 2**64).
    """
    if len(key) != 32:
        raise ValueError("ChaCha20 key must be 32 bytes")
    if len(nonce16) != 16:
        raise ValueError("ChaCha20 nonce must be 16 bytes")
    kw = np.frombuffer(key, dtype="<u4")
    nw = np.frombuffer(nonce16, dtype="<u4")
    ctr0 = int(nw[0]) | (int(nw[1]) << 32)
    ctr = (ctr0 + first_block + np.arange(nblocks, dtype=np.uint64).astype(object)) % (1 << 64)
    ctr = np.array(ctr, dtype=np.uint64)
    init = np.empty((16, nblocks), dtype=np.uint32)
    init[0:4] = _SIGMA[:, None]
    init[4:12] = kw[:, None]
    init[12] = (ctr & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    init[13] = (ctr >> np.uint64(32)).astype(np.uint32)
    init[14] = nw[2]
    init[15] = nw[3]
    s = [init[i].copy() for i in range(16)]
    with np.errstate(over="ignore"):
        for _ in range(10):
            _qr(s, 0, 4, 8, 12); _qr(s, 1, 5, 9, 13); _qr(s, 2, 6, 10, 14); _qr(s, 3, 7, 11, 15)
            _qr(s, 0, 5, 10, 15); _qr(s, 1, 6, 11, 12); _qr(s, 2, 7, 8, 13); _qr(s, 3, 4, 9, 14)
        out = np.stack([s[i] + init[i] for i in range(16)], axis=1)
    return out


def chacha20_keystream(key: bytes, nonce16: bytes, nbytes: int) -> np.ndarray:
    """First ``nbytes`` keystream bytes (uint8) for (key, nonce16)."""
    nblocks = (nbytes + 63) // 64
    words = chacha20_blocks(key, nonce16, nblocks)
    return words.astype("<u4").view(np.uint8).reshape(-1)[:nbytes].copy()


def chacha20_keystream_lib(key: bytes, nonce16: bytes, nbytes: int) -> np.ndarray:
    """Same keystream through the reference's own library (gs_insert.py:45-47)."""
    from cryptography.hazmat.backends import default_backend
    from cryptography.hazmat.primitives.ciphers import Cipher, algorithms

    enc = Cipher(algorithms.ChaCha20(key, nonce16), mode=None, backend=default_backend()).encryptor()
    return np.frombuffer(enc.update(bytes(nbytes)) + enc.finalize(), dtype=np.uint8).copy()


# --------------------------------------------------------------------------
# framing / key resolution
# --------------------------------------------------------------------------
def choose_watermark_length(total_blocks_needed: int) -> int:
    """nodes.py:26-49."""
    for thr, L in ((1024 * 32, 1024), (512 * 32, 512), (256 * 32, 256), (128 * 32, 128), (64 * 32, 64)):
        if total_blocks_needed >= thr:
            return L
    return 32


def pad_message(message, n_bytes: int) -> bytes:
    """gs_insert.py:9-16 / nodes.py:68-74: UTF-8, zero right-pad or truncate.

    ``message`` may be ``bytes`` (used verbatim, then padded) or anything ``str()`` accepts.
    An empty message means "random watermark" in the reference (os.urandom); the oracle
    requires the caller to pass the random bytes explicitly so results are reproducible.
    """
    mb = message if isinstance(message, (bytes, bytearray)) else str(message).encode()
    if len(mb) == 0:
        raise ValueError("oracle needs an explicit message (reference would draw os.urandom)")
    return bytes(mb[:n_bytes]) + b"\x00" * max(0, n_bytes - len(mb))


def frame_message(message, n_elems: int, l_bits: int = 256, use_repeat: bool = False):
    """Return (k, s_d): the padded message and the plaintext tiled over the latent.

    gs_insert.py:23 (k * 64), nodes.py:76-87 (k * (N // L) + zero tail),
    v1.5.2:29-47 (use_repeat: 8-byte message repeated 4x before tiling).
    """
    if l_bits % 8:
        raise ValueError("message length must be a whole number of bytes")
    if use_repeat:
        if l_bits % 32:
            raise ValueError("use_repeat needs a message length divisible by 4 bytes")
        k = pad_message(message, l_bits // 32) * 4
    else:
        k = pad_message(message, l_bits // 8)
    repeats = n_elems // l_bits
    s_d = k * repeats
    s_d += b"\x00" * ((n_elems + 7) // 8 - len(s_d))  # nodes.py:85-87, only the consumed part
    return k, s_d


def resolve_key_nonce(key_hex: str, nonce_hex: str):
    """gs_insert.py:27-42.  Empty key means os.urandom in the reference -> caller's job here."""
    if not key_hex:
        raise ValueError("oracle needs an explicit key (reference would draw os.urandom)")
    key = bytes.fromhex(key_hex)
    nonce = bytes.fromhex(nonce_hex) if nonce_hex else bytes.fromhex(key_hex[16:48])
    return key, nonce


# --------------------------------------------------------------------------
# embed
# --------------------------------------------------------------------------
def bucket_bits(s_d: bytes, key: bytes, nonce16: bytes, keystream=chacha20_keystream) -> np.ndarray:
    """y bit per latent element: MSB-first bits of  s_d XOR keystream  (gs_insert.py:45-49,58-60).

    ``keystream`` selects the numpy restatement (default) or ``chacha20_keystream_lib`` (the reference's
    own library call; used by bench.py's CPU baseline so the baseline is not slowed by numpy ChaCha)."""
    ks = keystream(key, nonce16, len(s_d))
    m = np.frombuffer(s_d, dtype=np.uint8) ^ ks
    return np.unpackbits(m)  # default bitorder='big' == format(byte, '08b')


def embed_from_uniform(y: np.ndarray, u: np.ndarray) -> np.ndarray:
    """gs_insert.py:64 with l = 1: z = norm.ppf((u + y) / 2), float64, same shape as u."""
    y = np.asarray(y, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    with np.errstate(divide="ignore"):
        return ndtri((u + y) / 2.0)


def embed(message, key: bytes, nonce16: bytes, u: np.ndarray, l_bits: int = 256,
          use_repeat: bool = False, keystream=chacha20_keystream) -> np.ndarray:
    """Whole embed for one latent; ``u`` is the flat float64 uniform stream (length N)."""
    n = int(np.asarray(u).size)
    _, s_d = frame_message(message, n, l_bits, use_repeat)
    y = bucket_bits(s_d, key, nonce16, keystream)[:n]               # the reference loop stops at N (nodes.py:122-123)
    return embed_from_uniform(y, np.asarray(u).reshape(-1))


def embed_scalar(message, key: bytes, nonce16: bytes, u, l_bits: int = 256) -> list:
    """Statement-by-statement loop form of gs_insert.py:23-66 (small N only)."""
    n = len(u)
    _, s_d = frame_message(message, n, l_bits)
    m = (np.frombuffer(s_d, dtype=np.uint8) ^ chacha20_keystream_lib(key, nonce16, len(s_d))).tobytes()
    m_bits = "".join(format(b, "08b") for b in m)
    out = []
    for i in range(n):
        y = int(m_bits[i], 2)
        out.append(float(ndtri((float(u[i]) + y) / 2)))
    return out


# --------------------------------------------------------------------------
# the product's counter-based uniform source (restated so the oracle can be fed
# the identical u).  Philox4x32-R (Salmon et al., SC'11); known answers are for R = 10.
# --------------------------------------------------------------------------
_PHILOX_M0 = np.uint64(0xD2511F53)
_PHILOX_M1 = np.uint64(0xCD9E8D57)
_PHILOX_W0 = 0x9E3779B9
_PHILOX_W1 = 0xBB67AE85


def philox4x32(ctr: np.ndarray, key, rounds: int = 10) -> np.ndarray:
    """ctr: (n, 4) uint32; key: (k0, k1).  Returns (n, 4) uint32."""
    c = [ctr[:, i].astype(np.uint64) for i in range(4)]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(rounds):
        p0 = _PHILOX_M0 * c[0]
        p1 = _PHILOX_M1 * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ np.uint64(k0), p1 & mask,
             (p0 >> np.uint64(32)) ^ c[3] ^ np.uint64(k1), p0 & mask]
        k0 = (k0 + _PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + _PHILOX_W1) & 0xFFFFFFFF
    return np.stack(c, axis=1).astype(np.uint32)


GSWM_TILE = 16384          # elements per tile (32 ChaCha blocks) -- the kernels' unit of work
GSWM_PHILOX_ROUNDS = 7           # csrc/gswm_math.cuh default (Philox4x32-7, the fewest Crush-resistant rounds)


def _gswm_counters(offset: int, latent_index: int, tiles: int, tile, s_, tid, call: int) -> np.ndarray:
    """Philox counters of "gswm uniforms v4" for arrays tile / s_ / tid (uint64) and one call index:
        T = (latent_index * tiles_per_latent + tile) * 4 + s        (< 2^54)
        ctr = (offset_lo, tid | (T & 0xFFFFFF) << 8, offset_hi, (T >> 24) << 2 | call)
    The lane index sits in word 1, which the first round does not multiply (csrc/gswm_math.cuh: philox4x32_v4_calls3)."""
    T = (np.uint64(latent_index * tiles) + tile) * np.uint64(4) + s_
    ctr = np.empty((T.size, 4), dtype=np.uint32)
    ctr[:, 0] = offset & 0xFFFFFFFF
    ctr[:, 1] = (tid | ((T & np.uint64(0xFFFFFF)) << np.uint64(8))).astype(np.uint32)
    ctr[:, 2] = (offset >> 32) & 0xFFFFFFFF
    ctr[:, 3] = ((((T >> np.uint64(24)) << np.uint64(2)) | np.uint64(call)) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    return ctr


def gswm_uniform_ints(seed: int, offset: int, latent_index: int, n_elems: int,
                      rounds: int = GSWM_PHILOX_ROUNDS) -> np.ndarray:
    """The 23-bit integer m of every element of one latent ("gswm uniforms v4", csrc/gswm_math.cuh).

    The latent is cut into tiles of 16384 elements; a tile into 4 super-iterations of 256 lanes; lane `tid`
    of super-iteration `s` owns the four float4 (16 elements) at within-tile float4 indices
    (4s + k) * 256 + tid, k = 0..3.  Three calls W_c = Philox4x32(ctr(T, tid, c), key = seed), c = 0..2 (counters:
    _gswm_counters) feed the 16 elements: float4 k < 3, element j: m = W_k[j] >> 9; float4 3, element j:
    m = (W_0[j] & 0xFF) | (W_1[j] & 0xFF) << 8 | (W_2[j] & 0x7F) << 16.
    """
    tiles = (n_elems + GSWM_TILE - 1) // GSWM_TILE
    t, s_, tid = np.meshgrid(np.arange(tiles, dtype=np.uint64), np.arange(4, dtype=np.uint64),
                             np.arange(256, dtype=np.uint64), indexing="ij")
    t, s_, tid = t.reshape(-1), s_.reshape(-1), tid.reshape(-1)
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    w = [philox4x32(_gswm_counters(offset, latent_index, tiles, t, s_, tid, c), key, rounds) for c in range(3)]   # (tiles*1024, 4)
    m = np.empty((tiles, 4, 4, 256, 4), dtype=np.uint32)            # [tile][s][k][tid][j]
    for k in range(3):
        m[:, :, k] = (w[k] >> np.uint32(9)).reshape(tiles, 4, 256, 4)
    low = (w[0] & np.uint32(0xFF)) | ((w[1] & np.uint32(0xFF)) << np.uint32(8)) | ((w[2] & np.uint32(0x7F)) << np.uint32(16))
    m[:, :, 3] = low.reshape(tiles, 4, 256, 4)
    # element index within tile = 4 * ((4s + k) * 256 + tid) + j  ->  C order of [s][k][tid][j]
    return m.reshape(-1)[:n_elems]


GSWM_TOP_CELL = (1 << 23) - 1   # the outermost grid cell, subdivided since "gswm uniforms v3"


def gswm_top_cell_words(seed: int, offset: int, latent_index: int, n_elems: int, elems: np.ndarray,
                        rounds: int = GSWM_PHILOX_ROUNDS) -> np.ndarray:
    """The 32 refinement bits of elements `elems` (flat indices into one latent) should they fall in the outermost cell:
    word j of Philox4x32(ctr(T, tid, call = 3), key = (seed_lo, seed_hi + k)), where the element is number j of float4 k
    of super-iteration T, lane tid (see gswm_uniform_ints)."""
    elems = np.asarray(elems, dtype=np.int64).reshape(-1)
    tiles = (n_elems + GSWM_TILE - 1) // GSWM_TILE
    tile, within = elems // GSWM_TILE, elems % GSWM_TILE
    f4, j = within // 4, within % 4
    tid, sk = f4 % 256, f4 // 256
    s_, k = sk // 4, sk % 4
    out = np.empty(elems.size, dtype=np.uint32)
    for i in range(elems.size):
        ctr = _gswm_counters(offset, latent_index, tiles, np.array([tile[i]], dtype=np.uint64), np.array([s_[i]], dtype=np.uint64),
                             np.array([tid[i]], dtype=np.uint64), 3)
        key = (seed & 0xFFFFFFFF, ((seed >> 32) + int(k[i])) & 0xFFFFFFFF)
        out[i] = philox4x32(ctr, key, rounds)[0, int(j[i])]
    return out


def gswm_uniforms(seed: int, offset: int, latent_index: int, y: np.ndarray,
                  rounds: int = GSWM_PHILOX_ROUNDS) -> np.ndarray:
    """float64 u in (0,1) for every element of one latent, given its bucket bits y ("gswm uniforms v4"):
    v = (m + 1/2) 2^-23;  u = v where y == 1, u = 1 - v where y == 0 (both exact in float64).  An element in the
    outermost cell m = 2^23 - 1 is refined by 28 more bits: 1 - v = (m2 + 1/2) 2^-51, m2 = refinement word >> 4."""
    y = np.asarray(y).reshape(-1)
    m = gswm_uniform_ints(seed, offset, latent_index, y.size, rounds)
    one_minus_v = (np.float64(1 << 23) - m.astype(np.float64) - 0.5) * 2.0 ** -23
    top = np.flatnonzero(m == GSWM_TOP_CELL)
    if top.size:
        w = gswm_top_cell_words(seed, offset, latent_index, y.size, top, rounds)
        one_minus_v[top] = ((w >> np.uint32(4)).astype(np.float64) + 0.5) * 2.0 ** -51
    return np.where(y == 1, 1.0 - one_minus_v, one_minus_v)


def embed_gswm(message, key: bytes, nonce16: bytes, seed: int, offset: int, latent_index: int, n_elems: int,
               l_bits: int = 256, rounds: int = GSWM_PHILOX_ROUNDS) -> np.ndarray:
    """What gswm_embed must produce for one latent: the reference formula fed the product's uniforms."""
    _, s_d = frame_message(message, n_elems, l_bits)
    y = bucket_bits(s_d, key, nonce16)[:n_elems]
    return embed_from_uniform(y, gswm_uniforms(seed, offset, latent_index, y, rounds))


# --------------------------------------------------------------------------
# extract
# --------------------------------------------------------------------------
def quantise(z: np.ndarray) -> np.ndarray:
    """extract.py:82-84 with l = 1: int(norm.cdf(z) * 2) per element, C order, uint8.

    Raises ValueError where the reference raises (NaN, or cdf*2 rounding to 2).
    """
    zf = np.asarray(z).astype(np.float64).reshape(-1)
    if np.isnan(zf).any():
        raise ValueError("cannot convert float NaN to integer")  # int(nan) in the reference
    q = (ndtr(zf) * 2.0).astype(np.int64)
    if (q >= 2).any():
        raise ValueError("invalid literal for int() with base 2")  # extract.py:86
    return q.astype(np.uint8)


def vote_counts(z: np.ndarray, key: bytes, nonce16: bytes, l_bits: int, keystream=chacha20_keystream) -> np.ndarray:
    """count_1 per message position (extract.py:86-98): uint32 [L]."""
    bits = quantise(z)
    n = bits.size
    if n % 8 or n % l_bits:
        raise ValueError("latent size must be a multiple of 8 and of the message length")
    m = np.packbits(bits)
    s_d = m ^ keystream(key, nonce16, m.size)
    all_bits = np.unpackbits(s_d)
    return all_bits.reshape(-1, l_bits).sum(axis=0).astype(np.uint32)


def vote_counts_batch(z: np.ndarray, key: bytes, nonce16: bytes, l_bits: int, keystream=chacha20_keystream) -> np.ndarray:
    """vote_counts for a batch [B, N] that shares one key / nonce: the keystream is produced once (extract.py:77-78 builds
    one cipher per call; the bytes are the same), everything else is the per-latent arithmetic of extract.py:82-98."""
    z = np.asarray(z)
    b, n = z.shape[0], int(np.prod(z.shape[1:]))
    bits = np.stack([quantise(z[i].reshape(-1)) for i in range(b)])
    ks = np.unpackbits(np.frombuffer(bytes(keystream(key, nonce16, (n + 7) // 8)), dtype=np.uint8))[:n]
    return (bits ^ ks[None, :]).reshape(b, n // l_bits, l_bits).sum(axis=1).astype(np.uint32)


def recover_message_bits(z: np.ndarray, key: bytes, nonce16: bytes, l_bits: int,
                         keystream=chacha20_keystream) -> np.ndarray:
    """extract.py:97-99: strict majority over the R = N / L copies; uint8 [L]."""
    counts = vote_counts(z, key, nonce16, l_bits, keystream)
    r = np.asarray(z).size // l_bits
    return (counts > r / 2).astype(np.uint8)


def recover_message(z: np.ndarray, key: bytes, nonce16: bytes, l_bits: int) -> str:
    """Return the '0'/'1' string extract.recover_exactracted_message returns."""
    return "".join("1" if b else "0" for b in recover_message_bits(z, key, nonce16, l_bits))


def recover_message_scalar(z, key: bytes, nonce16: bytes, l_bits: int) -> str:
    """Statement-by-statement loop form of extract.py:72-101 (small N only)."""
    bits = [int(float(ndtr(float(v))) * 2) for v in np.asarray(z, dtype=np.float64).reshape(-1)]
    m = bytes(int("".join(str(b) for b in bits[i:i + 8]), 2) for i in range(0, len(bits), 8))
    s_d = (np.frombuffer(m, dtype=np.uint8) ^ chacha20_keystream_lib(key, nonce16, len(m))).tobytes()
    all_bits = "".join("{:08b}".format(b) for b in s_d)
    segments = [all_bits[i:i + l_bits] for i in range(0, len(all_bits), l_bits)]
    out = ""
    for i in range(l_bits):
        count_1 = sum(seg[i] == "1" for seg in segments)
        out += "1" if count_1 > len(segments) / 2 else "0"
    return out


def calculate_bit_accuracy(original_message_hex: str, extracted_message_bin: str):
    """extract.py:103-110."""
    original = bin(int(original_message_hex, 16))[2:].zfill(len(original_message_hex) * 4)
    n = min(len(original), len(extracted_message_bin))
    original = original[:n]
    matching = sum(1 for a, b in zip(original, extracted_message_bin[:n]) if a == b)
    return original, matching / n


def bits_to_bytes(bits: np.ndarray) -> bytes:
    return np.packbits(np.asarray(bits, dtype=np.uint8)).tobytes()
