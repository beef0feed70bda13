#!/usr/bin/env python3
"""Generate tests/golden/golden_<tag>.json.gz by RUNNING THE REFERENCE ITSELF.

Run once in the build container (the only place /root/reference exists):

    python tests/golden/make_golden.py

For each shipped parameter set it imports the reference's golden model
(/root/reference/new_reference/cg_ntt.py, cg_ntt_8butterfly.py -- imported by
path, never copied), monkey-patches its module globals N and Q (they are read
at call time, cg_ntt.py:36,79), and records inputs/outputs of cg_ntt, cg_intt
and nwc_poly_mult for:
  * the seeds and KATs used by the reference's own tests
    (new_reference/test_cg_ntt.py:44-103, test_cg_ntt_8butterfly.py:49-118),
  * the LCG polynomials make_poly(1), make_poly(2) of the C++ benchmark,
  * edge vectors (zero, impulse, all-ones, all q-1, x^(N-1) * x wrap-around).
It also records the checksums printed by the reference's C++ code (via
oracle/_ref, built from the unmodified sources), the sha256 of the shipped
rtl/twiddle_*.hex tables, and the verbose log of one cg_ntt call.
Nothing in tests/ reads /root/reference at run time; they read these files.
"""
import gzip
import hashlib
import importlib.util
import json
import os
import random
import sys

REF = os.environ.get("TNTT_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ntt_oracle import PARAMS, make_poly  # noqa: E402  (only for the LCG inputs)
from oracle.cpu_ref import RefLib, build_ref  # noqa: E402

HEX = {
    "dilithium": ("twiddle_forward.hex", "twiddle_inverse.hex"),
    "n1024_24": ("twiddle_forward_1024.hex", "twiddle_inverse_1024.hex"),
    "n4096_24": ("twiddle_forward_4096.hex", "twiddle_inverse_4096.hex"),
    "n4096_60": ("twiddle_forward_4096_60bit.hex", "twiddle_inverse_4096_60bit.hex"),
}


def load_reference_modules():
    sys.path.insert(0, os.path.join(REF, "new_reference"))
    mods = {}
    for name in ("cg_ntt", "cg_ntt_8butterfly"):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, "new_reference", name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
    return mods["cg_ntt"], mods["cg_ntt_8butterfly"]


def sha256(path):
    with open(path, "rb") as fh:
        return hashlib.sha256(fh.read()).hexdigest()


def main():
    build_ref(REF)
    ref, ref8 = load_reference_modules()
    for tag, p in PARAMS.items():
        n, q, psi = p["n"], p["q"], p["psi"]
        ref.N, ref.Q = n, q            # looked up at call time (cg_ntt.py:36,79)
        ref8.N, ref8.Q = n, q
        omega = pow(psi, 2, q)
        g = {"tag": tag, "n": n, "q": q, "psi": psi, "omega": omega, "cases": {}}
        cases = g["cases"]

        def polymul_case(name, a, b):
            cases[name] = {"kind": "polymul", "a": a, "b": b, "c": ref.nwc_poly_mult(a, b, psi)}

        def ntt_case(name, a):
            fwd = ref.cg_ntt(a, omega, q)
            assert ref.cg_intt(fwd, omega, q) == [x % q for x in a]   # test_cg_ntt.py:44-52
            cases[name] = {"kind": "ntt", "a": a, "fwd": fwd, "intt_of_a": ref.cg_intt(a, omega, q)}

        # --- reference test seeds (test_cg_ntt.py:45,93; test_cg_ntt_8butterfly.py:50,61,109)
        random.seed(0)
        ntt_case("seed0_identity", [random.randrange(q) for _ in range(n)])
        random.seed(1)
        a = [random.randrange(q) for _ in range(n)]
        b = [random.randrange(q) for _ in range(n)]
        polymul_case("seed1_random", a, b)
        random.seed(2)
        ntt_case("seed2_identity8", [random.randrange(q) for _ in range(n)])
        random.seed(3)
        ntt_case("seed3_match8", [random.randrange(q) for _ in range(n)])
        random.seed(4)
        a = [random.randrange(q) for _ in range(n)]
        b = [random.randrange(q) for _ in range(n)]
        polymul_case("seed4_random8", a, b)
        assert ref8.nwc_poly_mult_8butterfly(a, b, psi) == cases["seed4_random8"]["c"]
        # --- KATs (test_cg_ntt.py:55-89)
        polymul_case("kat_123_456", [1, 2, 3] + [0] * (n - 3), [4, 5, 6] + [0] * (n - 3))
        polymul_case("kat_123_51", [1, 2, 3] + [0] * (n - 3), [5, 1] + [0] * (n - 2))
        # (1+5x+x^2)(5+x), test/cocotb_tests/test_ntt_inverse.py:273-275
        polymul_case("kat_151_51", [1, 5, 1] + [0] * (n - 3), [5, 1] + [0] * (n - 2))
        # --- edges
        zero = [0] * n
        imp = [1] + [0] * (n - 1)
        ones = [1] * n
        top = [q - 1] * n
        xn1 = [0] * (n - 1) + [1]
        x1 = [0, 1] + [0] * (n - 2)
        polymul_case("edge_zero", zero, top)
        polymul_case("edge_wrap", xn1, x1)
        polymul_case("edge_top", top, top)
        polymul_case("edge_one", imp, top)
        ntt_case("edge_impulse", imp)
        ntt_case("edge_ones", ones)
        ntt_case("edge_top_ntt", top)
        # --- LCG polynomials of the C++ benchmark
        a, b = make_poly(tag, 1), make_poly(tag, 2)
        polymul_case("lcg_1_2", a, b)
        ntt_case("lcg_1", a)
        tw = [pow(psi, i, q) for i in range(n)]
        cases["lcg_1_twisted_fwd"] = {"kind": "fwd_twist", "a": a,
                                      "fwd": ref.cg_ntt([x * w % q for x, w in zip(a, tw)], omega, q)}
        # --- unreduced / negative inputs are accepted
        if n == 256:
            random.seed(77)
            raw = [random.randrange(-q, 3 * q) for _ in range(n)]
            cases["unreduced"] = {"kind": "ntt_raw", "a": raw, "fwd": ref.cg_ntt(raw, omega, q)}
            log = []
            random.seed(0)
            a0 = [random.randrange(q) for _ in range(n)]
            ref.cg_ntt(a0, omega, q, verbose=True, log_fn=log.append)
            g["verbose_log_seed0"] = log
            log8 = []
            ref8.cg_ntt_8butterfly(a0, omega, q, verbose=True, log_fn=log8.append)
            g["verbose_log8_seed0"] = log8
        # --- the reference's C++ code on the same LCG inputs
        lib = RefLib(tag, "scalar")
        assert list(map(int, lib.make_poly(1))) == a
        out = lib.polymul(lib.make_poly(1), lib.make_poly(2))
        assert list(map(int, out)) == cases["lcg_1_2"]["c"], "C++ reference != Python reference"
        assert list(map(int, lib.schoolbook(lib.make_poly(1), lib.make_poly(2)))) == cases["lcg_1_2"]["c"]
        fwd = lib.forward(lib.make_poly(1))
        assert list(map(int, fwd)) == cases["lcg_1_twisted_fwd"]["fwd"]
        g["cpp_checksums"] = {"forward_ntt_checksum": lib.checksum(fwd), "checksum": lib.checksum(out)}
        for simd in ("avx2", "avx512"):
            other = RefLib(tag, simd)
            assert (other.polymul(other.make_poly(1), other.make_poly(2)) == out).all()
        # --- shipped twiddle tables
        f_hex, i_hex = HEX[tag]
        g["hex_sha256"] = {"forward": sha256(os.path.join(REF, "rtl", f_hex)),
                           "inverse": sha256(os.path.join(REF, "rtl", i_hex))}
        g["hex_names"] = {"forward": f_hex, "inverse": i_hex}
        with open(os.path.join(REF, "rtl", f_hex)) as fh:
            g["hex_head_forward"] = [fh.readline().strip() for _ in range(4)]

        if n > 256:   # keep the large-N fixtures small: the reference-test seeds, one KAT, the edges that wrap
            keep = ("seed0_identity", "seed1_random", "kat_123_456", "edge_wrap", "edge_top", "edge_impulse",
                    "lcg_1_2", "lcg_1", "lcg_1_twisted_fwd")
            g["cases"] = {k: v for k, v in cases.items() if k in keep}
        path = os.path.join(HERE, f"golden_{tag}.json.gz")
        with gzip.open(path, "wt", compresslevel=9) as fh:
            json.dump(g, fh, separators=(",", ":"))
        print(tag, "cases:", len(g["cases"]), "bytes:", os.path.getsize(path), g["cpp_checksums"])

    # n=4 worked example, test/refs/fast_ntt_negacyclic_convolution.py:161-214
    ref.N, ref.Q = 4, 7681
    small = {"n": 4, "q": 7681, "psi": 1925, "a": [1, 2, 3, 4], "b": [5, 6, 7, 8],
             "c": ref.nwc_poly_mult([1, 2, 3, 4], [5, 6, 7, 8], 1925)}
    assert small["c"] == [7625, 7645, 2, 60]
    with open(os.path.join(HERE, "golden_n4.json"), "w") as fh:
        json.dump(small, fh)


if __name__ == "__main__":
    main()
