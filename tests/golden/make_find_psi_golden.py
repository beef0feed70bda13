#!/usr/bin/env python3
"""Answers of the reference's own scripts/find_psi.py:9-44 (run in the build container, where /root/reference
exists) for its three built-in parameter sets and the four shipped rings -> tests/golden/golden_find_psi.json.
null = the script's search range [2, 10000) holds no primitive 2N-th root."""
import contextlib
import importlib.util
import io
import json
import os

spec = importlib.util.spec_from_file_location("ref_find_psi", "/root/reference/scripts/find_psi.py")
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)

CASES = [(256, 7681), (256, 8380417), (512, 12289), (1024, 8380417), (4096, 8380417),
         (4096, (1 << 60) - (1 << 14) + 1), (8, 17), (16, 97), (2048, 12289)]
out = []
for n, q in CASES:
    with contextlib.redirect_stdout(io.StringIO()):
        psi = mod.find_psi(n, q)
    out.append({"n": n, "q": q, "psi": psi})
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_find_psi.json")
with open(path, "w") as fh:
    json.dump({"source": "scripts/find_psi.py find_psi(n, q, max_search=10000)", "cases": out}, fh, indent=1)
    fh.write("\n")
print(open(path).read())
