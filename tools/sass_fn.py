#!/usr/bin/env python
"""Print the SASS of one function of a .so / .o / .cubin (first function whose mangled name contains the pattern),
one instruction per line without the encoding column.  Usage: python tools/sass_fn.py lib.so '<pattern>'"""
import re
import subprocess
import sys

sass = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
fn = [b for b in sass.split("Function : ") if sys.argv[2] in b.split("\n")[0]][0]
print(fn.split("\n")[0])
for ln in fn.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
    if m:
        print(f"{m.group(1)}  {m.group(2).strip()}")
