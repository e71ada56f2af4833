#!/usr/bin/env python
"""BASELINE config[0], timed on the UNMODIFIED reference: one SD-2.1 latent through
gs_insert.gs_watermark_init_noise (gs_insert.py:8-75), then extract.recover_exactracted_message +
calculate_bit_accuracy (extract.py:72-110) on that latent; one core, >= 3 repeats.

Runs in the BUILD container only (the GPU box has no /root/reference); writes profiles/r02_verbatim_reference.json,
which bench.py quotes in cpu_baseline.note.  The reference modules are imported exactly as tests/golden/make_golden.py
imports them (same sys.modules stubs for diffusers / matplotlib, which the codec functions never touch)."""
import importlib.util
import json
import os
import platform
import sys
import tempfile
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def main():
    os.sched_setaffinity(0, {sorted(os.sched_getaffinity(0))[0]})         # one core
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    mg._install_stubs()
    sys.path.insert(0, REF)
    gs_insert = mg._load(os.path.join(REF, "gs_insert.py"), "ref_gs_insert")
    extract = mg._load(os.path.join(REF, "extract.py"), "ref_extract")
    import torch

    opt = types.SimpleNamespace(key_hex=mg.KEY_HEX, nonce_hex=mg.NONCE_HEX)
    args = types.SimpleNamespace(key=bytes.fromhex(mg.KEY_HEX), nonce=bytes.fromhex(mg.NONCE_HEX), l=1, message_length=256)
    msg_hex = (b"lthero" + bytes(26)).hex()
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())                     # the reference appends ./info_data.txt
    rows = []
    try:
        for r in range(reps + 1):                    # the first pass pays the scipy / cryptography imports: not counted
            np.random.seed(1000 + r)
            t0 = time.perf_counter()
            z = gs_insert.gs_watermark_init_noise(opt, "lthero")
            t1 = time.perf_counter()
            lat = torch.tensor(z).float().reshape(1, 4, 64, 64)        # what the caller hands on (README.md:112)
            bits = extract.recover_exactracted_message(lat, args)
            _, acc = extract.calculate_bit_accuracy(msg_hex, bits)
            t2 = time.perf_counter()
            assert acc == 1.0
            if r:
                rows.append({"embed_s": t1 - t0, "extract_s": t2 - t1, "bit_accuracy": acc})
    finally:
        os.chdir(cwd)
    emb = float(np.median([x["embed_s"] for x in rows]))
    ext = float(np.median([x["extract_s"] for x in rows]))
    import scipy
    import cryptography
    out = {"what": "BASELINE config[0]: verbatim reference, one 4x64x64 latent, gs_insert.gs_watermark_init_noise then "
                   "extract.recover_exactracted_message + calculate_bit_accuracy, default key/nonce, message 'lthero'",
           "cores": 1, "repeats": reps, "embed_s_median": emb, "extract_s_median": ext, "pair_s_median": emb + ext,
           "pairs_per_s": 1.0 / (emb + ext), "runs": rows,
           "host": {"cpu": platform.processor() or platform.machine(), "python": platform.python_version(),
                    "numpy": np.__version__, "scipy": scipy.__version__, "cryptography": cryptography.__version__},
           "where": "build container (the GPU box has no /root/reference)"}
    path = os.path.join(ROOT, "profiles", "r02_verbatim_reference.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: out[k] for k in ("embed_s_median", "extract_s_median", "pairs_per_s")}))


if __name__ == "__main__":
    main()
