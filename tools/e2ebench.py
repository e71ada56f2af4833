#!/usr/bin/env python
"""Where the host-buffer (e2e) path spends its time: the pipe legs alone and together, against the same
work done by the kernels addressing pinned host memory directly (zero-copy over PCIe), and the plain
copy-engine ceilings.  Run under gpurun:  python tools/e2ebench.py"""
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "a-watermark-for-diffusion-models_b200"))
import gswm  # noqa: E402
from gswm.codec import _DeviceJob  # noqa: E402

B, shape, L = int(os.environ.get("EB_B", 4096)), (4, 64, 64), 256
n = int(np.prod(shape))
dev = torch.device("cuda:0")
lib = gswm._lib.lib()
km = gswm.KeyMaterial.make(bytes.fromhex(gswm.DEFAULT_KEY_HEX), bytes.fromhex(gswm.DEFAULT_NONCE_HEX),
                           gswm.pad_message("lthero", L // 8), L)
z = gswm.embed_batch(B, shape, km, 0x5EED, 0, 0, dev)
h_in = (z + 0.325 * torch.randn_like(z)).cpu().pin_memory()
h_out = torch.empty((B, *shape), dtype=torch.float32).pin_memory()
chunk = int(os.environ.get("EB_CHUNK", 256))
pe = gswm.HostPipe(0, max_elems=n, chunk_latents=chunk)
px = gswm.HostPipe(0, max_elems=n, chunk_latents=chunk)


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best, tot = 1e9, 0.0
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        best = min(best, dt)
        tot += dt
    return {"ms_best": round(best * 1e3, 3), "ms_mean": round(tot / reps * 1e3, 3), "GBps_best": round(B * n * 4 / best / 1e9, 1)}


def both(f, g):
    def run():
        t = threading.Thread(target=f)
        t.start()
        g()
        t.join()
    return run


res = {"B": B, "chunk": chunk}
res["pipe_embed_alone"] = timeit(lambda: pe.embed(h_out, km, 0x5EED))
res["pipe_extract_alone"] = timeit(lambda: px.extract(h_in, km))
res["pipe_both"] = timeit(both(lambda: pe.embed(h_out, km, 0x5EED), lambda: px.extract(h_in, km)))

# zero-copy: the kernels read / write the pinned host buffers through their device-visible addresses
dj = _DeviceJob(km, B, n, dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
msgs = torch.empty((B, L // 8), dtype=torch.uint8, device=dev)
matched = torch.empty((B,), dtype=torch.int32, device=dev)
counters = torch.zeros(4, dtype=torch.int64, device=dev)
ws2 = None   # no entry point needs scratch memory (gswm_workspace_bytes == 0)


def zc_embed(sync=True):
    gswm._lib.check(lib.gswm_embed(C.byref(dj.job), 0x5EED, 0, 0, h_out.data_ptr(), dj.ws_ptr, s1.cuda_stream), "embed")
    if sync:
        s1.synchronize()


def zc_extract(sync=True):
    gswm._lib.check(lib.gswm_extract(C.byref(dj.job), h_in.data_ptr(), 0, msgs.data_ptr(), None, matched.data_ptr(),
                                     counters.data_ptr(), None, s2.cuda_stream), "extract")
    if sync:
        s2.synchronize()


def zc_both():
    zc_embed(False)
    zc_extract(False)
    s1.synchronize()
    s2.synchronize()


try:
    res["zerocopy_embed_alone"] = timeit(zc_embed)
    ok = bool(torch.equal(h_out[:64], z[:64].cpu()) and torch.equal(h_out[-64:], z[-64:].cpu()))
    res["zerocopy_embed_ok"] = ok
    res["zerocopy_extract_alone"] = timeit(zc_extract)
    counters.zero_()
    zc_extract()
    res["zerocopy_extract_counters"] = counters.cpu().tolist()
    res["zerocopy_both"] = timeit(zc_both)
except Exception as e:  # noqa: BLE001
    res["zerocopy_error"] = repr(e)

# copy-engine ceilings
d_a = torch.empty(B * n, dtype=torch.float32, device=dev)
d_b = torch.empty(B * n, dtype=torch.float32, device=dev)


def ce(h2d, d2h):
    def run():
        if h2d:
            with torch.cuda.stream(s1):
                d_a.copy_(h_in.view(-1), non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.view(-1).copy_(d_b, non_blocking=True)
        s1.synchronize()
        s2.synchronize()
    return run


res["ce_h2d"] = timeit(ce(True, False))
res["ce_d2h"] = timeit(ce(False, True))
res["ce_both"] = timeit(ce(True, True))
print(json.dumps(res, indent=1))
