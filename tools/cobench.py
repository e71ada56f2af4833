#!/usr/bin/env python
"""Co-scheduling experiment: embed on one stream and extract on another, each with a capped number of resident CTAs
per SM (env GSWM_EMBED_CTAS_PER_SM / GSWM_EXTRACT_CTAS_PER_SM), against the same two kernels back to back on one
stream.  Usage: python tools/cobench.py lib.so [embed_ctas extract_ctas]...   (run under gpurun)"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "a-watermark-for-diffusion-models_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from gswm import _lib  # noqa: E402
from kbench import load  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    B, n, Lb = int(os.environ.get("KB_B", 4096)), 16384, 256
    key = bytes.fromhex("5822ff9cce6772f714192f43863f6bad1bf54b78326973897e6b66c3186b77a7")
    nonce = bytes.fromhex("05072fd1c2265f6f2e2a4080a2bfbdd8")
    msg = b"lthero" + bytes(26)
    flat = torch.from_numpy(np.frombuffer(key + nonce + msg, np.uint8).copy()).to(dev)
    job = _lib.Job(B, n, Lb, 0, flat.data_ptr(), flat.data_ptr() + 32, flat.data_ptr() + 48)
    z = torch.empty((B, n), dtype=torch.float32, device=dev)
    zn = torch.empty((B, n), dtype=torch.float32, device=dev)
    msgs = torch.empty((B, 32), dtype=torch.uint8, device=dev)
    matched = torch.empty((B,), dtype=torch.int32, device=dev)
    counters = torch.zeros(4, dtype=torch.int64, device=dev)
    L = load(sys.argv[1])
    ws = torch.empty(max(16, L.gswm_workspace_bytes(C.byref(job))), dtype=torch.uint8, device=dev)
    ws2 = torch.empty_like(ws)
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    reps = int(os.environ.get("KB_REPS", 200))

    def embed(st):
        assert L.gswm_embed(C.byref(job), 0x5EED, 0, 0, z.data_ptr(), ws.data_ptr(), st.cuda_stream) == 0

    def extract(st):
        assert L.gswm_extract(C.byref(job), zn.data_ptr(), 0, msgs.data_ptr(), None, matched.data_ptr(), counters.data_ptr(),
                              ws2.data_ptr(), st.cuda_stream) == 0

    embed(sa)
    torch.cuda.synchronize()
    zn.copy_(z + 0.325 * torch.randn_like(z))
    torch.cuda.synchronize()
    pairs = [(int(a), int(b)) for a, b in zip(sys.argv[2::2], sys.argv[3::2])] or [(0, 0)]
    for ec, xc in pairs:
        for k, v in (("GSWM_EMBED_CTAS_PER_SM", ec), ("GSWM_EXTRACT_CTAS_PER_SM", xc)):
            if v:
                os.environ[k] = str(v)
            else:
                os.environ.pop(k, None)
        res = {"lib": os.path.basename(sys.argv[1]), "embed_ctas": ec, "extract_ctas": xc}
        for mode in ("serial", "concurrent", "embed_only", "extract_only"):
            best = 1e9
            for _ in range(3):
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True)
                ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(sa)
                sb.wait_event(e0)
                for _ in range(reps):
                    if mode == "serial":
                        embed(sa), extract(sa)
                    elif mode == "concurrent":
                        embed(sa), extract(sb)
                    elif mode == "embed_only":
                        embed(sa)
                    else:
                        extract(sb)
                ea.record(sa)
                eb.record(sb)
                torch.cuda.synchronize()
                best = min(best, max(e0.elapsed_time(ea), e0.elapsed_time(eb)) * 1e3 / reps)
            res[mode + "_us"] = round(best, 2)
        counters.zero_()
        extract(sb)
        torch.cuda.synchronize()
        res["exact"] = counters.cpu().tolist()
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
