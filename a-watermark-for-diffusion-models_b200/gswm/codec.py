"""Batch API of the Gaussian-Shading codec on device tensors (and, through :class:`HostPipe`, on host
buffers).  PyTorch is used for device memory and streams only; all arithmetic happens in libgswm.so.

The per-latent semantics are those of the reference:
  embed    gs_insert.gs_watermark_init_noise (gs_insert.py:8-66), nodes.gs_watermark_init_noise (nodes.py:51-123)
  extract  extract.recover_exactracted_message + calculate_bit_accuracy (extract.py:72-110)
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence, Union

import numpy as np
import torch

from . import _lib
from ._lib import GSWM_BF16, GSWM_F16, GSWM_F32, GSWM_F64, Job

DEFAULT_KEY_HEX = "5822ff9cce6772f714192f43863f6bad1bf54b78326973897e6b66c3186b77a7"  # README.md:61
DEFAULT_NONCE_HEX = "05072fd1c2265f6f2e2a4080a2bfbdd8"  # README.md:67

_DTYPE_CODE = {torch.float32: GSWM_F32, torch.float16: GSWM_F16, torch.bfloat16: GSWM_BF16, torch.float64: GSWM_F64}
BytesLike = Union[bytes, bytearray, np.ndarray, torch.Tensor]


# ----------------------------------------------------------------------------- host-side framing
def choose_watermark_length(total_blocks_needed: int) -> int:
    """Message length policy of the ComfyUI node (nodes.py:26-49)."""
    for thr, bits in ((1024 * 32, 1024), (512 * 32, 512), (256 * 32, 256), (128 * 32, 128), (64 * 32, 64)):
        if total_blocks_needed >= thr:
            return bits
    return 32


def pad_message(message, n_bytes: int, use_repeat: bool = False) -> bytes:
    """The padded / truncated watermark ``k`` (gs_insert.py:9-20; nodes.py:68-76; v1.5.2:29-47).

    ``message`` is a str (UTF-8 encoded like the reference's ``str(message).encode()``) or raw bytes.
    An empty message draws ``os.urandom`` exactly as the reference does.  ``use_repeat`` is the webui
    option: a quarter-length message repeated four times.
    """
    import os

    unit = n_bytes // 4 if use_repeat else n_bytes
    if isinstance(message, (bytes, bytearray)):
        mb = bytes(message)
    else:
        mb = str(message).encode() if message else b""
    k = (mb[:unit] + b"\x00" * max(0, unit - len(mb))) if mb else os.urandom(unit)
    return k * 4 if use_repeat else k


def resolve_key_nonce(key_hex: str, nonce_hex: str):
    """key / nonce bytes from the hex strings, with the reference's fallbacks (gs_insert.py:27-42):
    empty nonce -> key_hex[16:48]; empty key -> both random.  Malformed hex raises ValueError as
    bytes.fromhex does in the reference; wrong lengths raise ValueError like `cryptography` does."""
    import os

    if key_hex:
        key = bytes.fromhex(key_hex)
        nonce = bytes.fromhex(nonce_hex) if nonce_hex else bytes.fromhex(key_hex[16:48])
    else:
        key, nonce = os.urandom(32), os.urandom(16)
    if len(key) != 32:
        raise ValueError("Invalid key size (%d) for ChaCha20." % (len(key) * 8))
    if len(nonce) != 16:
        raise ValueError("nonce must be 128-bits (16 bytes)")
    return key, nonce


def _as_u8_rows(x: BytesLike, row_bytes: int, what: str) -> np.ndarray:
    """bytes or array -> contiguous uint8 [rows, row_bytes] numpy array."""
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    if isinstance(x, (bytes, bytearray)):
        a = np.frombuffer(bytes(x), dtype=np.uint8)
    else:
        a = np.ascontiguousarray(x, dtype=np.uint8)
    if a.size % row_bytes:
        raise ValueError(f"{what}: expected a multiple of {row_bytes} bytes, got {a.size}")
    return a.reshape(-1, row_bytes)


@dataclass
class KeyMaterial:
    """Key / nonce / message bytes for a job: shared (one row each) or one row per latent."""
    keys: np.ndarray      # uint8 [rows, 32]
    nonces: np.ndarray    # uint8 [rows, 16]
    msgs: Optional[np.ndarray]  # uint8 [rows, msg_bits/8] or None (extract without reference message)
    msg_bits: int

    @classmethod
    def make(cls, key: BytesLike, nonce: BytesLike, msg: Optional[BytesLike], msg_bits: int) -> "KeyMaterial":
        """``msg_bits``: embedding needs a positive multiple of 32; extraction takes any positive length that divides
        the latent (extract.py:195 ``--message_length`` is an arbitrary integer).  Message rows are (msg_bits + 7) // 8 bytes."""
        if msg_bits <= 0:
            raise ValueError("message length must be positive")
        k = _as_u8_rows(key, 32, "key")
        n = _as_u8_rows(nonce, 16, "nonce")
        m = None if msg is None else _as_u8_rows(msg, (msg_bits + 7) // 8, "message")
        rows = {k.shape[0], n.shape[0]} | ({m.shape[0]} if m is not None else set())
        big = max(rows)
        if rows - {1, big}:
            raise ValueError("key / nonce / message row counts disagree")

        def bc(a):
            return a if a is None or a.shape[0] == big else np.ascontiguousarray(np.broadcast_to(a, (big, a.shape[1])))

        return cls(bc(k), bc(n), bc(m), msg_bits)

    @property
    def rows(self) -> int:
        return self.keys.shape[0]

    @property
    def per_latent(self) -> bool:
        return self.rows > 1


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


_KM_CACHE = {}            # (device index, key material bytes) -> device tensor; small shared-key material only
_KM_CACHE_MAX = 1024


def _upload_key_material(packed: np.ndarray, device: torch.device) -> torch.Tensor:
    """[keys | nonces | msgs] on the device.  Shared key material (a few dozen bytes, the reference's usual case: one
    --key_hex / --nonce_hex / --message per run) is uploaded ONCE per device and reused by every later call with the same
    bytes, so a call site that is invoked per latent enqueues nothing but its kernel -- no synchronous pageable copy in
    front of it, which also keeps programmatic dependent launch working between consecutive calls."""
    if packed.nbytes > 4096:
        return torch.from_numpy(packed).to(device, non_blocking=False)
    ck = (device.index, packed.tobytes())
    t = _KM_CACHE.get(ck)
    if t is None:
        if len(_KM_CACHE) >= _KM_CACHE_MAX:      # kernels on any stream may still be reading an entry: drain before dropping them
            torch.cuda.synchronize(device)
            _KM_CACHE.clear()
        t = torch.from_numpy(packed).to(device, non_blocking=False)
        _KM_CACHE[ck] = t
    return t


class _DeviceJob:
    """Key material on the device + the ctypes job describing it (keeps the tensor alive)."""

    def __init__(self, km: KeyMaterial, n_latents: int, n_elems: int, device: torch.device, keys_in_flight: bool = False):
        if km.per_latent and km.rows != n_latents:
            raise ValueError(f"per-latent key material has {km.rows} rows for {n_latents} latents")
        packed = [km.keys.reshape(-1), km.nonces.reshape(-1)]
        if km.msgs is not None:
            packed.append(km.msgs.reshape(-1))
        # one buffer for all three arrays (keys and nonces are 32 / 16 bytes per row: the message segment stays 4-byte aligned)
        flat = _upload_key_material(np.concatenate(packed), device)
        self.flat = flat
        o1 = km.keys.size
        o2 = o1 + km.nonces.size
        flags = (_lib.JOB_PER_LATENT if km.per_latent else 0) | (_lib.JOB_KEYS_IN_FLIGHT if keys_in_flight else 0)
        self.job = Job(n_latents, n_elems, km.msg_bits, flags,
                       flat.data_ptr(), flat.data_ptr() + o1, (flat.data_ptr() + o2) if km.msgs is not None else None)


def _device(device) -> torch.device:
    d = torch.device(device)
    if d.type != "cuda":
        raise ValueError("gswm runs on CUDA devices only (there is no CPU path)")
    if d.index is None:
        d = torch.device("cuda", torch.cuda.current_device())
    return d


def _n_elems(shape: Sequence[int]) -> int:
    n = int(np.prod(shape))
    if n <= 0 or n % 4:
        raise ValueError(f"latent size {tuple(shape)} must be a positive multiple of 4 elements")
    return n


# ----------------------------------------------------------------------------- device-tensor API
def chacha20_keystream(keys: BytesLike, nonces: BytesLike, n_bytes: int, device="cuda") -> torch.Tensor:
    """Keystream bytes [rows, n_bytes] (uint8, device) -- what the reference gets from
    ``Cipher(algorithms.ChaCha20(key, nonce)).encryptor().update(bytes(n))`` (gs_insert.py:45-47)."""
    dev = _device(device)
    k = _as_u8_rows(keys, 32, "key")
    n = _as_u8_rows(nonces, 16, "nonce")
    if k.shape[0] != n.shape[0]:
        raise ValueError("need one nonce per key")
    padded = (n_bytes + 63) // 64 * 64
    with torch.cuda.device(dev):
        dk = torch.from_numpy(k.copy()).to(dev)
        dn = torch.from_numpy(n.copy()).to(dev)
        out = torch.empty((k.shape[0], padded), dtype=torch.uint8, device=dev)
        _lib.check(_lib.lib().gswm_chacha20_keystream(dk.data_ptr(), dn.data_ptr(), k.shape[0], padded, out.data_ptr(),
                                                      _stream_ptr(dev)), "gswm_chacha20_keystream")
    return out[:, :n_bytes]


def embed_batch(n_latents: int, latent_shape: Sequence[int], km: KeyMaterial, seed: int, offset: int = 0,
                first_latent: int = 0, device="cuda", out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Watermarked initial noise [n_latents, *latent_shape] fp32 on ``device`` with the in-kernel
    uniform source (Philox4x32-7 keyed by ``seed``; global latent index ``first_latent + b``)."""
    dev = _device(device)
    n = _n_elems(latent_shape)
    with torch.cuda.device(dev):
        dj = _DeviceJob(km, n_latents, n, dev)
        if out is None:
            out = torch.empty((n_latents, *latent_shape), dtype=torch.float32, device=dev)
        elif out.dtype != torch.float32 or out.numel() != n_latents * n or not out.is_contiguous() or out.device != dev:
            raise ValueError("out must be a contiguous fp32 tensor of n_latents * n_elems elements on the device")
        if n_latents == 0:                       # an empty batch is valid and launches nothing
            return out
        _lib.check(_lib.lib().gswm_embed(C.byref(dj.job), seed & (2 ** 64 - 1), offset & (2 ** 64 - 1), first_latent,
                                         out.data_ptr(), _stream_ptr(dev)), "gswm_embed")
        # per-latent key material is freed when dj goes out of scope; the caching allocator keeps the block tied to this
        # stream, so reuse is stream-ordered.
    return out


def embed_batch_injected(u: torch.Tensor, latent_shape: Sequence[int], km: KeyMaterial, n_latents: Optional[int] = None,
                         out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """Embed with injected float64 uniforms ``u`` ([n_latents, n_elems] or [n_elems] shared), evaluated in
    float64 exactly as z = norm.ppf((u + y) / 2) (gs_insert.py:64)."""
    n = _n_elems(latent_shape)
    if u.dtype != torch.float64 or not u.is_cuda:
        raise ValueError("u must be a float64 CUDA tensor")
    dev = u.device
    u = u.contiguous()
    shared_u = u.numel() == n and (n_latents is not None and n_latents != 1)
    if n_latents is None:
        n_latents = u.numel() // n
    if not shared_u and u.numel() != n_latents * n:
        raise ValueError("u has the wrong number of elements")
    with torch.cuda.device(dev):
        dj = _DeviceJob(km, n_latents, n, dev)
        out = torch.empty((n_latents, *latent_shape), dtype=out_dtype, device=dev)
        _lib.check(_lib.lib().gswm_embed_injected(C.byref(dj.job), u.data_ptr(), 0 if shared_u else 1, out.data_ptr(),
                                                  _DTYPE_CODE[out_dtype], _stream_ptr(dev)),
                   "gswm_embed_injected")
    return out


def _seed_args(seeds, n: int, dev: torch.device):
    """(device tensor or None, seed0) for the MT19937 entry points.  numpy's legacy seeding takes integers in [0, 2^32)
    (anything else raises ValueError there too)."""
    if isinstance(seeds, (int, np.integer)):
        if not 0 <= int(seeds) <= 0xFFFFFFFF:
            raise ValueError("Seed must be between 0 and 2**32 - 1")
        return None, int(seeds)
    arr = np.ascontiguousarray(seeds, dtype=np.int64).reshape(-1)
    if arr.size != n:
        raise ValueError(f"need one seed per stream ({n}), got {arr.size}")
    if arr.size and (arr.min() < 0 or arr.max() > 0xFFFFFFFF):
        raise ValueError("Seed must be between 0 and 2**32 - 1")
    return torch.from_numpy(arr.astype(np.uint32).view(np.int32)).to(dev), 0


def mt19937_uniform(seeds, n_each: int, n_streams: Optional[int] = None, device="cuda") -> torch.Tensor:
    """[n_streams, n_each] float64 on the device: row s is ``np.random.RandomState(seed_s).uniform(size=n_each)`` bit for
    bit (MT19937, init_genrand seeding, 53-bit doubles).  ``seeds``: one int (stream s uses seed + s) or one per stream."""
    dev = _device(device)
    if n_streams is None:
        n_streams = 1 if isinstance(seeds, (int, np.integer)) else len(seeds)
    with torch.cuda.device(dev):
        d_seeds, seed0 = _seed_args(seeds, n_streams, dev)
        out = torch.empty((n_streams, n_each), dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().gswm_mt19937_uniform(d_seeds.data_ptr() if d_seeds is not None else None, seed0, n_streams,
                                                   n_each, out.data_ptr(), _stream_ptr(dev)), "gswm_mt19937_uniform")
    return out


def embed_batch_mt19937(seeds, n_latents: int, latent_shape: Sequence[int], km: KeyMaterial,
                        out_dtype: torch.dtype = torch.float32, device="cuda") -> torch.Tensor:
    """The reference's SEEDED embed for a batch, entirely on the device: latent b draws its uniforms from
    ``np.random.RandomState(seed_b)`` element by element (nodes.py:52-53,114-117; v1.5.2:27,72-75) -- generated by the
    MT19937 kernel, nothing uploaded -- and z = norm.ppf((u + y) / 2) is evaluated in float64 (gs_insert.py:64).
    ``seeds``: one int (latent b uses seed + b) or one per latent."""
    dev = _device(device)
    n = _n_elems(latent_shape)
    if km.msg_bits % 32:
        raise ValueError("embedding needs a message length that is a multiple of 32 bits")
    if 0 < n_latents * n * 8 <= (64 << 20):
        # small batches (every call of the seeded drop-ins is ONE latent): generator kernel into a scratch tensor, then the
        # injected-uniform kernel, which spreads a latent over 16 CTAs -- the fused kernel has the stream's one CTA evaluate
        # all of its quantiles as well.  Same uniforms, same arithmetic, identical output; 66 against 104 us for one SD-2.1
        # latent, 149 against 283 us for one SDXL latent (tools/mt19937_routes.py).
        return embed_batch_injected(mt19937_uniform(seeds, n, n_latents, dev), latent_shape, km, n_latents, out_dtype)
    with torch.cuda.device(dev):
        d_seeds, seed0 = _seed_args(seeds, n_latents, dev)
        dj = _DeviceJob(km, n_latents, n, dev)
        out = torch.empty((n_latents, *latent_shape), dtype=out_dtype, device=dev)
        if n_latents == 0:
            return out
        _lib.check(_lib.lib().gswm_embed_mt19937(C.byref(dj.job), d_seeds.data_ptr() if d_seeds is not None else None, seed0,
                                                 out.data_ptr(), _DTYPE_CODE[out_dtype], _stream_ptr(dev)),
                   "gswm_embed_mt19937")
    return out


@dataclass
class ExtractResult:
    messages: torch.Tensor            # uint8 [B, msg_bits/8], MSB-first packed decoded bits
    counts: Optional[torch.Tensor]    # uint16 [B, msg_bits] count_1 per position
    matched: Optional[torch.Tensor]   # int32 [B] bits equal to the reference message
    counters: torch.Tensor            # int64 [6]: matched_bits, total_bits, exact_msgs, total_msgs, nan_latents, range_latents
    flags: Optional[torch.Tensor] = None   # uint8 [B]: FLAG_NAN / FLAG_RANGE -- inputs the reference raises on (extract.py:83,86)
    msg_bits: int = 0
    reduced: Optional[torch.Tensor] = None # int64 [6]: the counters summed over the ranks of a Comm (extract_batch(comm=...))

    def bit_strings(self):
        """Decoded messages as the '0'/'1' strings extract.recover_exactracted_message returns."""
        bits = np.unpackbits(self.messages.cpu().numpy(), axis=1)
        if self.msg_bits:
            bits = bits[:, :self.msg_bits]          # a length that is not a whole number of bytes ends inside the last byte
        return ["".join("1" if b else "0" for b in row) for row in bits]

    def bit_accuracy(self) -> float:
        c = self.counters.cpu()
        return float(c[0]) / float(c[1]) if int(c[1]) else float("nan")    # nan: no latent was scored


def extract_batch(z: torch.Tensor, km: KeyMaterial, want_counts: bool = False,
                  counters: Optional[torch.Tensor] = None, comm: Optional["Comm"] = None) -> ExtractResult:
    """Decode a batch of inverted latents ``z`` [B, ...] (fp32 / fp16 / bf16 / fp64, CUDA).  ``flags`` marks the
    latents the reference would refuse (a NaN, or an element >= 8.2924); they are decoded all the same with
    bit = (z >= threshold).  With ``comm`` the cross-GPU sum of the accumulated counters is fused into the same
    kernel (gswm_extract_allreduce) and returned as ``result.reduced``."""
    if not z.is_cuda:
        raise ValueError("z must be a CUDA tensor (there is no CPU path)")
    if z.dtype not in (torch.float32, torch.float16, torch.bfloat16, torch.float64):
        raise ValueError(f"unsupported latent dtype {z.dtype}")
    dev = z.device
    z = z.contiguous()
    b = z.shape[0]
    n = _n_elems(z.shape[1:])
    if n % km.msg_bits:
        raise ValueError("message length must divide the latent size")
    row = (km.msg_bits + 7) // 8
    with torch.cuda.device(dev):
        dj = _DeviceJob(km, b, n, dev)
        msgs = torch.empty((b, row), dtype=torch.uint8, device=dev)
        cnt = torch.empty((b, km.msg_bits), dtype=torch.uint16, device=dev) if want_counts else None
        matched = torch.empty((b,), dtype=torch.int32, device=dev) if km.msgs is not None else None
        flags = torch.empty((b,), dtype=torch.uint8, device=dev)
        if counters is None:
            counters = torch.zeros((_lib.N_COUNTERS,), dtype=torch.int64, device=dev)
        res = ExtractResult(msgs, cnt, matched, counters, flags, km.msg_bits)
        if b == 0 and comm is None:              # an empty batch is valid and launches nothing (with a communicator the
            return res                           # rank still joins the exchange: the peers wait for it)
        args = (C.byref(dj.job), z.data_ptr(), _DTYPE_CODE[z.dtype], msgs.data_ptr(),
                cnt.data_ptr() if cnt is not None else None, matched.data_ptr() if matched is not None else None,
                flags.data_ptr(), counters.data_ptr())
        if comm is None:
            _lib.check(_lib.lib().gswm_extract(*args, _stream_ptr(dev)), "gswm_extract")
        else:
            res.reduced = torch.empty((_lib.N_COUNTERS,), dtype=torch.int64, device=dev)
            _lib.check(_lib.lib().gswm_extract_allreduce(*args, comm.handle, res.reduced.data_ptr(), _stream_ptr(dev)),
                       "gswm_extract_allreduce")
    return res


_side_streams = {}


def embed_extract_batch(n_latents: int, latent_shape: Sequence[int], km_embed: KeyMaterial, seed: int,
                        z_in: torch.Tensor, km_extract: Optional[KeyMaterial] = None, offset: int = 0,
                        first_latent: int = 0, want_counts: bool = False, out: Optional[torch.Tensor] = None):
    """One service step: embed ``n_latents`` new latents AND decode the batch ``z_in`` of inverted latents, the two
    kernels co-scheduled on two CUDA streams.  Embed is bound by the SMs' FMA pipes, extract by HBM, so side by side
    they finish in about the time the pair's bytes need at HBM speed (92 us instead of 110 us one after the other for
    2 x 4096 SD-2.1 latents).  Results are those of :func:`embed_batch` and :func:`extract_batch`; both are ordered
    after the caller's stream on entry and before it on return."""
    dev = z_in.device
    if dev.type != "cuda":
        raise ValueError("z_in must be a CUDA tensor (there is no CPU path)")
    cur = torch.cuda.current_stream(dev)
    side = _side_streams.get(dev.index)
    if side is None:
        side = _side_streams[dev.index] = torch.cuda.Stream(dev)
    fork = torch.cuda.Event()
    fork.record(cur)
    side.wait_event(fork)
    with torch.cuda.stream(side):
        res = extract_batch(z_in, km_extract or km_embed, want_counts)
    z_in.record_stream(side)
    z_out = embed_batch(n_latents, latent_shape, km_embed, seed, offset, first_latent, dev, out)
    join = torch.cuda.Event()
    join.record(side)
    cur.wait_event(join)
    for t in (res.messages, res.counts, res.matched, res.counters, res.flags):
        if t is not None:
            t.record_stream(cur)
    return z_out, res


# ----------------------------------------------------------------------------- host-buffer API
class HostPipe:
    """gswm_pipe_*: batches in HOST memory (numpy / CPU torch), chunked and overlapped over PCIe."""

    def __init__(self, device=0, max_elems: int = 65536, chunk_latents: int = 256):
        self._p = C.c_void_p()
        dev = torch.device(device) if not isinstance(device, int) else torch.device("cuda", device)
        self.device_index = dev.index or 0
        _lib.check(_lib.lib().gswm_pipe_create(C.byref(self._p), self.device_index, max_elems, chunk_latents),
                   "gswm_pipe_create")

    def close(self):
        if self._p:
            _lib.lib().gswm_pipe_destroy(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    @staticmethod
    def _host_job(km: KeyMaterial, n_latents: int, n_elems: int):
        if km.per_latent and km.rows != n_latents:
            raise ValueError("per-latent key material row count != n_latents")
        keep = (np.ascontiguousarray(km.keys), np.ascontiguousarray(km.nonces),
                None if km.msgs is None else np.ascontiguousarray(km.msgs))
        job = Job(n_latents, n_elems, km.msg_bits, _lib.JOB_PER_LATENT if km.per_latent else 0, keep[0].ctypes.data, keep[1].ctypes.data,
                  keep[2].ctypes.data if keep[2] is not None else None)
        return job, keep

    def embed(self, out: Union[np.ndarray, torch.Tensor], km: KeyMaterial, seed: int, offset: int = 0,
              first_latent: int = 0):
        """Fill ``out`` (host fp32 [B, ...], ideally pinned) with watermarked noise."""
        t = torch.from_numpy(out) if isinstance(out, np.ndarray) else out
        if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError("out must be a contiguous fp32 host array")
        b = t.shape[0]
        n = _n_elems(t.shape[1:])
        if b == 0:
            return out
        job, keep = self._host_job(km, b, n)
        _lib.check(_lib.lib().gswm_pipe_embed(self._p, C.byref(job), seed & (2 ** 64 - 1), offset & (2 ** 64 - 1),
                                              first_latent, t.data_ptr()), "gswm_pipe_embed")
        del keep
        return out

    def embed_injected(self, u: np.ndarray, latent_shape: Sequence[int], km: KeyMaterial, n_latents: int = 1,
                       out_dtype=np.float32) -> np.ndarray:
        """Embed with injected float64 uniforms on the host; returns a numpy array [n_latents, *latent_shape]."""
        n = _n_elems(latent_shape)
        u = np.ascontiguousarray(u, dtype=np.float64)
        per = 0 if (u.size == n and n_latents != 1) else 1
        if per and u.size != n * n_latents:
            raise ValueError("u has the wrong number of elements")
        out = np.empty((n_latents, *latent_shape), dtype=out_dtype)
        job, keep = self._host_job(km, n_latents, n)
        code = GSWM_F32 if np.dtype(out_dtype) == np.float32 else GSWM_F64
        _lib.check(_lib.lib().gswm_pipe_embed_injected(self._p, C.byref(job), u.ctypes.data, per, out.ctypes.data, code),
                   "gswm_pipe_embed_injected")
        del keep
        return out

    def extract(self, z: Union[np.ndarray, torch.Tensor], km: KeyMaterial, want_counts: bool = False):
        """Decode host latents [B, ...]; returns (messages u8 [B, ceil(L/8)], counts u16 or None, matched i32 or None,
        counters i64[6], flags u8 [B]) as numpy arrays."""
        t = torch.from_numpy(z) if isinstance(z, np.ndarray) else z
        if t.is_cuda or t.dtype not in (torch.float32, torch.float16, torch.bfloat16, torch.float64):
            raise ValueError("z must be a host fp32 / fp16 / bf16 / fp64 array")
        t = t.contiguous()
        b = t.shape[0]
        n = _n_elems(t.shape[1:])
        job, keep = self._host_job(km, b, n)
        msgs = np.empty((b, (km.msg_bits + 7) // 8), dtype=np.uint8)
        cnt = np.empty((b, km.msg_bits), dtype=np.uint16) if want_counts else None
        matched = np.empty((b,), dtype=np.int32) if km.msgs is not None else None
        flags = np.zeros((b,), dtype=np.uint8)
        counters = np.zeros((_lib.N_COUNTERS,), dtype=np.int64)
        if b == 0:
            return msgs, cnt, matched, counters, flags
        _lib.check(_lib.lib().gswm_pipe_extract(self._p, C.byref(job), t.data_ptr(), _DTYPE_CODE[t.dtype], msgs.ctypes.data,
                                                cnt.ctypes.data if cnt is not None else None,
                                                matched.ctypes.data if matched is not None else None,
                                                flags.ctypes.data, counters.ctypes.data), "gswm_pipe_extract")
        del keep
        return msgs, cnt, matched, counters, flags


# ----------------------------------------------------------------------------- multi-GPU
class Comm:
    """gswm_comm: per-rank mailboxes mapped over NVLink (include/gswm.h, "Multi-GPU").  One instance per rank; the
    CUDA IPC handles are exchanged through ``torch.distributed`` (any initialised backend), which is used for that
    rendezvous only -- the all-reduce itself is a gswm kernel storing into the peers' memory."""

    def __init__(self, device=None, group=None):
        import torch.distributed as dist

        dev = _device(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
        self.device = dev
        if dist.is_available() and dist.is_initialized():
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        else:
            self.rank, self.world = 0, 1
        self._h = C.c_void_p()
        handle = (C.c_uint8 * _lib.COMM_HANDLE_BYTES)()
        _lib.check(_lib.lib().gswm_comm_create(C.byref(self._h), dev.index, self.rank, self.world, handle), "gswm_comm_create")
        if self.world > 1:
            mine = torch.tensor(list(handle), dtype=torch.uint8)
            backend = dist.get_backend(group)
            if backend == "nccl":
                mine = mine.to(dev)
            gathered = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(gathered, mine, group=group)
            blob = torch.cat([g.cpu() for g in gathered]).numpy().tobytes()
            buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
            _lib.check(_lib.lib().gswm_comm_connect(self._h, buf), "gswm_comm_connect")
            dist.barrier(group=group)            # every rank has mapped every mailbox before anyone publishes

    @property
    def handle(self):
        return self._h

    def allreduce_counters(self, counters: torch.Tensor) -> torch.Tensor:
        """Sum ``counters`` (int64, <= 8 values, on this rank's device) over all ranks, in place, on the current stream."""
        if counters.dtype != torch.int64 or not counters.is_cuda or not counters.is_contiguous():
            raise ValueError("counters must be a contiguous int64 CUDA tensor")
        _lib.check(_lib.lib().gswm_comm_allreduce_counters(self._h, counters.data_ptr(), counters.numel(),
                                                           _stream_ptr(counters.device)), "gswm_comm_allreduce_counters")
        return counters

    def status(self) -> int:
        return int(_lib.lib().gswm_comm_status(self._h))

    def close(self):
        if self._h:
            _lib.lib().gswm_comm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
