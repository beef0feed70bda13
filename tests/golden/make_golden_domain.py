#!/usr/bin/env python3
"""Outputs of the reference's own new_reference/cg_ntt.py (imported from /root/reference, build container only) for
arguments OUTSIDE the NTT-friendly domain: even / composite moduli, non-primitive roots, omega = 0.  The reference
never validates them (cg_ntt.py:29-92); these vectors pin the drop-in's literal path to the same numbers.
-> tests/golden/golden_domain.json"""
import importlib.util
import json
import os
import random

spec = importlib.util.spec_from_file_location("ref_cg_ntt", "/root/reference/new_reference/cg_ntt.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

CASES = [
    # n, q, omega, psi, note
    (16, 7681 * 3, 3, 5, "odd composite modulus"),
    (16, 1 << 20, 3, 7, "even modulus (power of two)"),
    (32, 2, 1, 1, "q = 2"),
    (8, 17, 4, 2, "prime q; omega of order 4, psi with psi^8 = 1 (not -1)"),
    (8, 17, 0, 0, "omega = 0, psi = 0"),
    (64, (2**30 - 35) * (2**29 - 3), 123456789, 987654321, "59-bit product of two primes"),
    (256, 8380416, 1753, 1753, "even 23-bit modulus next to the Dilithium prime"),
    (16, 97, 8, 19, "prime q, primitive psi (sanity: the fast and the literal path agree)"),
]
rng = random.Random(20261018)
out = []
for n, q, omega, psi, note in CASES:
    ref.N, ref.Q = n, q
    a = [rng.randrange(q) for _ in range(n)]
    b = [rng.randrange(q) for _ in range(n)]
    fwd = ref.cg_ntt(a, omega, q)
    out.append({"n": n, "q": q, "omega": omega, "psi": psi, "note": note, "a": a, "b": b,
                "cg_ntt": fwd, "cg_intt": ref.cg_intt(a, omega, q), "roundtrip": ref.cg_intt(fwd, omega, q),
                "nwc_poly_mult": ref.nwc_poly_mult(a, b, psi)})
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_domain.json")
with open(path, "w") as fh:
    json.dump({"source": "new_reference/cg_ntt.py cg_ntt / cg_intt / nwc_poly_mult with module globals N, Q set per case",
               "cases": out}, fh)
    fh.write("\n")
print(path, os.path.getsize(path), "bytes;", len(out), "cases")
