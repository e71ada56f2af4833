#!/usr/bin/env python
"""Kernel-variant timing harness: for each libgswm build given on the command line, time embed-only,
extract-only and alternating loops (B=4096 SD-2.1 latents, shared key) with CUDA events around the whole loop.
Usage: python tools/kbench.py variant1.so variant2.so ...   (run under gpurun)"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "a-watermark-for-diffusion-models_b200"))
from gswm import _lib  # noqa: E402


def load(path):
    L = C.CDLL(path)
    vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
    JP = C.POINTER(_lib.Job)
    L.gswm_embed.argtypes = [JP, u64, u64, i64, vp, vp]
    L.gswm_extract.argtypes = [JP, vp, i32, vp, vp, vp, vp, vp, vp]
    return L


def main():
    dev = torch.device("cuda:0")
    B, n, Lb = int(os.environ.get("KB_B", 4096)), int(os.environ.get("KB_N", 16384)), 256
    key = bytes.fromhex("5822ff9cce6772f714192f43863f6bad1bf54b78326973897e6b66c3186b77a7")
    nonce = bytes.fromhex("05072fd1c2265f6f2e2a4080a2bfbdd8")
    msg = b"lthero" + bytes(26)
    per = int(os.environ.get("KB_PER_LATENT", 0))
    zdt = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}[os.environ.get("KB_ZDTYPE", "f32")]
    zcode = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}[zdt]
    if per:
        rs = np.random.RandomState(2025)
        kb, nb, mb = rs.bytes(32 * B), rs.bytes(16 * B), rs.bytes(32 * B)
        flat = torch.from_numpy(np.frombuffer(kb + nb + mb, np.uint8).copy()).to(dev)
        job = _lib.Job(B, n, Lb, 1, flat.data_ptr(), flat.data_ptr() + 32 * B, flat.data_ptr() + 48 * B)
    else:
        flat = torch.from_numpy(np.frombuffer(key + nonce + msg, np.uint8).copy()).to(dev)
        job = _lib.Job(B, n, Lb, 0, flat.data_ptr(), flat.data_ptr() + 32, flat.data_ptr() + 48)
    z = torch.empty((B, n), dtype=torch.float32, device=dev)
    zn = torch.empty((B, n), dtype=zdt, device=dev)
    msgs = torch.empty((B, 32), dtype=torch.uint8, device=dev)
    matched = torch.empty((B,), dtype=torch.int32, device=dev)
    counters = torch.zeros(6, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    reps = int(os.environ.get("KB_REPS", 200))
    for path in sys.argv[1:]:
        L = load(path)
        def embed():
            rc = L.gswm_embed(C.byref(job), 0x5EED, 0, 0, z.data_ptr(), st)
            assert rc == 0, rc

        def extract():
            rc = L.gswm_extract(C.byref(job), zn.data_ptr(), zcode, msgs.data_ptr(), None, matched.data_ptr(), None,
                                counters.data_ptr(), st)
            assert rc == 0, rc

        embed()
        torch.cuda.synchronize()
        zn.copy_(z + 0.325 * torch.randn_like(z))
        res = {"lib": os.path.basename(path), "per_latent": per, "zdtype": str(zdt), "B": B, "n": n}
        legs = (("embed_us", embed), ("extract_us", extract), ("pair_us", lambda: (embed(), extract())))
        if os.environ.get("KB_ONLY_EMBED"):                     # tools/whatif.sh: diagnostic builds, embed timing only
            legs = legs[:1]
        for name, fn in legs:
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = 1e9
            for _ in range(3):
                e0.record()
                for _ in range(reps):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) * 1e3 / reps)
            res[name] = round(best, 2)
        bytes_ = B * n * 4
        res["embed_GBps"] = round(bytes_ / res["embed_us"] / 1e3, 1)
        if "extract_us" in res:
            counters.zero_()
            extract()
            torch.cuda.synchronize()
            res["exact"] = counters.cpu().tolist()
            res["extract_GBps"] = round(B * n * zn.element_size() / res["extract_us"] / 1e3, 1)
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
