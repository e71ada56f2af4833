#!/usr/bin/env python
"""Bit-identity of embed outputs across libgswm builds: every build given on the command line must write exactly the
latents the first one writes (SD-2.1, SDXL and a 3-tile shape; two batch sizes; a non-zero first_latent).
Usage: python tools/embed_identity.py base.so other.so ...   (run under gpurun)"""
import ctypes as C
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
    key = bytes.fromhex("5822ff9cce6772f714192f43863f6bad1bf54b78326973897e6b66c3186b77a7")
    nonce = bytes.fromhex("05072fd1c2265f6f2e2a4080a2bfbdd8")
    msg = b"lthero" + bytes(26)
    flat = torch.from_numpy(np.frombuffer(key + nonce + msg, np.uint8).copy()).to(dev)
    st = torch.cuda.current_stream().cuda_stream
    libs = [load(p) for p in sys.argv[1:]]
    ok = True
    for n in (16384, 65536, 49152, 32768):
        for B, first in ((1, 0), (7, 3), (300, 1 << 33), (4096, 12345)):
            job = _lib.Job(B, n, 256, 0, flat.data_ptr(), flat.data_ptr() + 32, flat.data_ptr() + 48)
            outs = []
            for L in libs:
                z = torch.full((B, n), float("nan"), dtype=torch.float32, device=dev)
                rc = L.gswm_embed(C.byref(job), 0x5EED, 7, first, z.data_ptr(), None, st)
                assert rc == 0, rc
                torch.cuda.synchronize()
                outs.append(z)
            for p, z in zip(sys.argv[2:], outs[1:]):
                same = bool(torch.equal(z.view(torch.int32), outs[0].view(torch.int32)))
                ok &= same
                print(f"n={n} B={B} first={first} {os.path.basename(p)}: {'identical' if same else 'DIFFERENT'}", flush=True)
    print("ALL IDENTICAL" if ok else "MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
