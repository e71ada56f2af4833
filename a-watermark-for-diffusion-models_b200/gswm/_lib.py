"""ctypes binding of libgswm.so (C ABI declared in include/gswm.h).

The library is built in-tree by :func:`build` (plain ``nvcc -gencode arch=compute_100a,code=sm_100a``;
no torch headers) and must exist for anything in this package to work: there is no CPU or PyTorch
fallback, a missing or unloadable library raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(os.path.dirname(_HERE), "csrc")
_INCLUDE = os.path.join(os.path.dirname(os.path.dirname(_HERE)), "include")
# GSWM_LIB: load another build of the same ABI instead (A/B runs of kernel variants through bench.py / the tests)
LIB_PATH = os.environ.get("GSWM_LIB") or os.path.join(_HERE, "libgswm.so")
SOURCES = ["gswm_kernels.cu", "gswm_mt19937.cu", "gswm_comm.cu", "gswm_pipe.cu", "gswm_microbench.cu"]

ABI_VERSION = 2
GSWM_F32, GSWM_F16, GSWM_BF16, GSWM_F64 = 0, 1, 2, 3
CTR_MATCHED_BITS, CTR_TOTAL_BITS, CTR_EXACT_MSGS, CTR_TOTAL_MSGS, CTR_NAN_LATENTS, CTR_RANGE_LATENTS, N_COUNTERS = range(7)
FLAG_NAN, FLAG_RANGE = 1, 2
JOB_PER_LATENT, JOB_KEYS_IN_FLIGHT = 1, 2
E_COMM = -6
ISSUE_KINDS = {"FFMA2": 0, "IMAD.WIDE": 1, "LOP3": 2, "MUFU.LG2": 3, "FFMA(imm)": 4, "IMAD.WIDE+LOP3": 5,
               "FFMA2|LOP3": 6, "FFMA|IMAD.WIDE": 7, "FFMA2|IMAD.WIDE": 8, "LOP3|IMAD.WIDE": 9}
COMM_HANDLE_BYTES, COMM_MAX_VALUES, COMM_MAX_RANKS = 64, 8, 32

EXPORTS = [
    "gswm_abi_version", "gswm_strerror", "gswm_chacha20_keystream", "gswm_embed", "gswm_embed_injected",
    "gswm_embed_mt19937", "gswm_mt19937_uniform", "gswm_extract", "gswm_extract_allreduce", "gswm_comm_create",
    "gswm_comm_connect", "gswm_comm_connect_local", "gswm_comm_destroy", "gswm_comm_status",
    "gswm_comm_allreduce_counters", "gswm_allreduce_counters", "gswm_pipe_create", "gswm_pipe_destroy",
    "gswm_pipe_embed", "gswm_pipe_embed_injected", "gswm_pipe_extract", "gswm_launch_count",
    "gswm_debug_bucket_quantile", "gswm_debug_norm_ppf", "gswm_debug_top_cell", "gswm_debug_philox4x32", "gswm_debug_issue_rate", "gswm_philox_rounds",
]


class GswmError(RuntimeError):
    def __init__(self, code: int, what: str):
        super().__init__(f"{what} failed: {code} ({strerror(code)})")
        self.code = code


class Job(C.Structure):
    """struct gswm_job / gswm_host_job (identical layout)."""
    _fields_ = [("n_latents", C.c_int64), ("n_elems", C.c_int64), ("msg_bits", C.c_int32),
                ("flags", C.c_int32), ("keys", C.c_void_p), ("nonces", C.c_void_p), ("msgs", C.c_void_p)]


def nvcc_command(out: str = LIB_PATH, extra=()):
    srcs = [os.path.join(_CSRC, s) for s in SOURCES]
    return ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
            "-Xcompiler", "-fPIC", "-shared", f"-I{_INCLUDE}", *extra, "-o", out, *srcs, "-ldl"]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libgswm.so for sm_100a (cross-compiles without a GPU).  Returns its path."""
    if os.environ.get("GSWM_LIB"):          # an explicitly chosen build is never rebuilt or overwritten
        return LIB_PATH
    srcs = [os.path.join(_CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cuh", ".inc", ".h"))]
    deps.append(os.path.join(_INCLUDE, "gswm.h"))
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    tmp = LIB_PATH + ".building"                # link into a scratch name, then rename: a reader never sees half a library
    cmd = nvcc_command(out=tmp, extra=("-Xptxas", "-v") if verbose else ())
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    os.replace(tmp, LIB_PATH)
    if verbose:
        print(res.stderr)
    return LIB_PATH


_lib = None
_lock = threading.Lock()


def lib() -> C.CDLL:
    """Load libgswm.so once; raise loudly when it is missing (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not built -- run `python __graft_entry__.py build` "
                              "(gswm has no CPU / PyTorch fallback)")
        L = C.CDLL(LIB_PATH)
        vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
        JP = C.POINTER(Job)
        L.gswm_abi_version.restype = C.c_int
        L.gswm_strerror.restype = C.c_char_p
        L.gswm_strerror.argtypes = [C.c_int]
        L.gswm_chacha20_keystream.argtypes = [vp, vp, i64, i64, vp, vp]
        L.gswm_embed.argtypes = [JP, u64, u64, i64, vp, vp]
        L.gswm_embed_injected.argtypes = [JP, vp, i32, vp, i32, vp]
        L.gswm_embed_mt19937.argtypes = [JP, vp, C.c_uint32, vp, i32, vp]
        L.gswm_mt19937_uniform.argtypes = [vp, C.c_uint32, i64, i64, vp, vp]
        L.gswm_extract.argtypes = [JP, vp, i32, vp, vp, vp, vp, vp, vp]
        L.gswm_extract_allreduce.argtypes = [JP, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp]
        L.gswm_comm_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, vp]
        L.gswm_comm_connect.argtypes = [vp, vp]
        L.gswm_comm_connect_local.argtypes = [C.POINTER(vp), C.c_int]
        L.gswm_comm_destroy.argtypes = [vp]
        L.gswm_comm_destroy.restype = None
        L.gswm_comm_status.argtypes = [vp]
        L.gswm_comm_allreduce_counters.argtypes = [vp, vp, i32, vp]
        L.gswm_allreduce_counters.argtypes = [vp, vp, i32, vp]
        L.gswm_pipe_create.argtypes = [C.POINTER(vp), C.c_int, i64, i64]
        L.gswm_pipe_destroy.argtypes = [vp]
        L.gswm_pipe_destroy.restype = None
        L.gswm_pipe_embed.argtypes = [vp, JP, u64, u64, i64, vp]
        L.gswm_pipe_embed_injected.argtypes = [vp, JP, vp, i32, vp, i32]
        L.gswm_pipe_extract.argtypes = [vp, JP, vp, i32, vp, vp, vp, vp, vp]
        L.gswm_launch_count.restype = i64
        L.gswm_philox_rounds.restype = C.c_int
        L.gswm_debug_bucket_quantile.argtypes = [vp, i64, i32, i32, vp, vp]
        L.gswm_debug_norm_ppf.argtypes = [vp, i64, vp, vp]
        L.gswm_debug_philox4x32.argtypes = [vp, i64, i32, vp, vp]
        L.gswm_debug_top_cell.argtypes = [vp, i64, vp, vp]
        L.gswm_debug_issue_rate.argtypes = [i32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.gswm_debug_issue_rate.restype = C.c_int
        for name in ("gswm_chacha20_keystream", "gswm_embed", "gswm_embed_injected", "gswm_embed_mt19937",
                     "gswm_mt19937_uniform", "gswm_extract", "gswm_extract_allreduce", "gswm_comm_create",
                     "gswm_comm_connect", "gswm_comm_connect_local", "gswm_comm_status", "gswm_comm_allreduce_counters",
                     "gswm_allreduce_counters", "gswm_pipe_create", "gswm_pipe_embed", "gswm_pipe_embed_injected",
                     "gswm_pipe_extract", "gswm_debug_bucket_quantile", "gswm_debug_norm_ppf", "gswm_debug_philox4x32",
                     "gswm_debug_top_cell"):
            getattr(L, name).restype = C.c_int
        if L.gswm_abi_version() != ABI_VERSION:
            raise ImportError(f"libgswm.so ABI version {L.gswm_abi_version()} != {ABI_VERSION}: rebuild (python __graft_entry__.py build)")
        _lib = L
        return L


def strerror(code: int) -> str:
    return lib().gswm_strerror(int(code)).decode()


def check(code: int, what: str) -> None:
    if code != 0:
        raise GswmError(code, what)


def launch_count() -> int:
    return int(lib().gswm_launch_count())
