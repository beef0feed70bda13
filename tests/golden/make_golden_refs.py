#!/usr/bin/env python3
"""Golden vectors for the reference's test/refs twins (ntt_forward_reference / ntt_inverse_reference).

Run in the build container only (it imports the reference from /root/reference/test):
    python tests/golden/make_golden_refs.py
Writes tests/golden/golden_refs.json.  Inputs deliberately include unreduced and negative integers,
because the twins reduce ``% q`` on entry (ntt_forward_reference.py:50)."""
import json
import os
import random
import sys

sys.path.insert(0, "/root/reference/test")
from refs.ntt_forward_reference import ntt_forward_reference  # noqa: E402
from refs.ntt_inverse_reference import ntt_inverse_reference  # noqa: E402

SETS = [(8, 8380417, pow(1239911, 32, 8380417)), (256, 8380417, 1239911), (1024, 8380417, 5548360),
        (256, (1 << 60) - (1 << 14) + 1, pow(431606828070683274, 16, (1 << 60) - (1 << 14) + 1))]
out = []
for n, q, psi in SETS:
    rng = random.Random(n * 31 + q % 1000)
    x = [rng.randrange(-q, 3 * q) for _ in range(n)]
    fwd = ntt_forward_reference(x, n, q, psi)
    inv = ntt_inverse_reference(x, n, q, psi)
    assert ntt_inverse_reference(fwd, n, q, psi) == [v % q for v in x]
    out.append({"n": n, "q": q, "psi": psi, "x": x, "forward": fwd, "inverse": inv})
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_refs.json")
with open(path, "w") as fh:
    json.dump(out, fh)
print("wrote", path, os.path.getsize(path), "bytes")
