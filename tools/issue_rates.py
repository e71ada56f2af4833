#!/usr/bin/env python
"""Issue-rate microbenchmarks of libgswm (gswm_debug_issue_rate): warp instructions per clock per SM sub-partition for the
embed kernel's instruction classes and for pairs of classes issued side by side (do the pipes overlap?).
Usage: python tools/issue_rates.py   (run under gpurun)"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "a-watermark-for-diffusion-models_b200"))
import gswm  # noqa: E402

lib = gswm._lib.lib()
out = {}
for name, kind in gswm._lib.ISSUE_KINDS.items():
    r, g = C.c_double(), C.c_double()
    rc = lib.gswm_debug_issue_rate(kind, C.byref(r), C.byref(g))
    out[name] = {"rc": rc, "warp_inst_per_clk_per_smsp": round(r.value, 4), "cycles_per_inst": round(1 / r.value, 3) if r.value else None,
                 "sm_ghz": round(g.value, 3)}
print(json.dumps(out, indent=1))
