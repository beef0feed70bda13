"""Pin the oracle (oracle/) against the reference's own vectors.  CPU only.

The golden files were produced by importing and running the reference
(tests/golden/make_golden.py); the checksums come from the reference's C++ code.
"""
import json
import os
import random

import numpy as np
import pytest

from conftest import GOLDEN_DIR, TAGS, load_golden
from oracle import ntt_oracle as O
from oracle.cpu_ref import COracle, RefLib, best_simd


@pytest.fixture(scope="module")
def coracle():
    return COracle()


def test_parameter_sets_are_valid():
    # psi is a primitive 2N-th root, Barrett constants of precompute_constants.py
    for tag, p in O.PARAMS.items():
        assert O.is_primitive_2n_root(p["psi"], p["n"], p["q"]), tag
    assert O.barrett_constants(8380417) == (23, 8396807)
    assert O.barrett_constants(O.PARAMS["n4096_60"]["q"]) == (60, 1152921504606863359)
    assert O.modinv(256, 8380417) == 8347681
    assert O.modinv(4096, O.PARAMS["n4096_60"]["q"]) == 1152640029630119941


def test_python_oracle_matches_golden(golden):
    n, q, psi, omega = golden["n"], golden["q"], golden["psi"], golden["omega"]
    for name, c in golden["cases"].items():
        if c["kind"] == "polymul":
            assert O.nwc_poly_mult(c["a"], c["b"], psi, q) == c["c"], name
        elif c["kind"] == "ntt":
            assert O.cg_ntt(c["a"], omega, q) == c["fwd"], name
            assert O.cg_intt(c["fwd"], omega, q) == [x % q for x in c["a"]], name
            assert O.cg_intt(c["a"], omega, q) == c["intt_of_a"], name
        elif c["kind"] == "fwd_twist":
            assert O.forward_negacyclic(c["a"], psi, q) == c["fwd"], name
        elif c["kind"] == "ntt_raw":
            assert O.cg_ntt(c["a"], omega, q) == c["fwd"], name


def test_c_oracle_matches_golden(golden, coracle):
    n, q, psi, omega = golden["n"], golden["q"], golden["psi"], golden["omega"]
    for name, c in golden["cases"].items():
        if c["kind"] == "polymul":
            assert coracle.nwc_poly_mult(c["a"], c["b"], psi, q).tolist() == c["c"], name
        elif c["kind"] == "ntt":
            assert coracle.cg_ntt(c["a"], omega, q).tolist() == c["fwd"], name
            assert coracle.cg_intt(c["a"], omega, q).tolist() == c["intt_of_a"], name


def test_checksums_match_reference_cpp(golden, coracle):
    tag, n, q, psi = golden["tag"], golden["n"], golden["q"], golden["psi"]
    a, b = O.make_poly(tag, 1), O.make_poly(tag, 2)
    assert a == golden["cases"]["lcg_1_2"]["a"] and b == golden["cases"]["lcg_1_2"]["b"]
    assert coracle.make_poly(1, n, q).tolist() == a
    want = golden["cpp_checksums"]
    assert O.checksum(tag, O.forward_negacyclic(a, psi, q)) == want["forward_ntt_checksum"]
    c = O.nwc_poly_mult(a, b, psi, q)
    assert O.checksum(tag, c) == want["checksum"]
    assert coracle.checksum(c, q) == want["checksum"]


SURVEY_CHECKSUMS = {  # SURVEY.md section 4 table / BASELINE.md section 2
    "dilithium": (16403698204383513489, 16424788039373839479),
    "n1024_24": (3555142461877891881, 15308795525113097448),
    "n4096_24": (2800297349529693940, 11303505593119465445),
    "n4096_60": (15678418584317678507, 2710933653778106521),
}


def test_golden_checksums_are_the_surveyed_ones(golden):
    fwd, mul = SURVEY_CHECKSUMS[golden["tag"]]
    assert golden["cpp_checksums"] == {"forward_ntt_checksum": fwd, "checksum": mul}


def test_kats():
    q, psi, n = 8380417, 1239911, 256
    assert O.nwc_poly_mult([1, 2, 3] + [0] * 253, [4, 5, 6] + [0] * 253, psi, q)[:5] == [4, 13, 28, 27, 18]
    assert O.nwc_poly_mult([1, 2, 3] + [0] * 253, [5, 1] + [0] * 254, psi, q)[:4] == [5, 11, 17, 3]
    assert O.nwc_poly_mult([1, 5, 1] + [0] * 253, [5, 1] + [0] * 254, psi, q)[:4] == [5, 26, 10, 1]
    wrap = O.nwc_poly_mult([0] * 255 + [1], [0, 1] + [0] * 254, psi, q)
    assert wrap == [q - 1] + [0] * 255
    omega = psi * psi % q
    assert O.cg_ntt([1] + [0] * 255, omega, q) == [1] * 256
    assert O.cg_ntt([1] * 256, omega, q) == [256] + [0] * 255
    with open(os.path.join(GOLDEN_DIR, "golden_n4.json")) as fh:
        g = json.load(fh)
    assert O.nwc_poly_mult(g["a"], g["b"], g["psi"], g["q"]) == g["c"] == [7625, 7645, 2, 60]


def test_cg_ntt_is_the_natural_order_dft():
    q, psi = 8380417, 1239911
    rng = random.Random(5)
    for n in (2, 4, 8, 32):
        omega = pow(psi, 2 * 256 // n, q)
        a = [rng.randrange(q) for _ in range(n)]
        assert O.cg_ntt(a, omega, q) == O.naive_dft(a, omega, q)


def test_polymul_equals_schoolbook(coracle):
    rng = random.Random(6)
    for tag in ("dilithium", "n1024_24"):
        p = O.PARAMS[tag]
        a = [rng.randrange(p["q"]) for _ in range(p["n"])]
        b = [rng.randrange(p["q"]) for _ in range(p["n"])]
        want = O.schoolbook_negacyclic(a, b, p["q"])
        assert O.nwc_poly_mult(a, b, p["psi"], p["q"]) == want
        assert coracle.schoolbook(a, b, p["q"]).tolist() == want


def test_barrett_matches_modulo():
    # scripts/precompute_constants.py:145-172 self-test, widened
    rng = random.Random(42)
    for q in (8380417, O.PARAMS["n4096_60"]["q"]):
        k, mu = O.barrett_constants(q)
        for _ in range(20000):
            a, b = rng.randrange(q), rng.randrange(q)
            assert O.barrett_reduce(a * b, q, k, mu) == a * b % q
        assert O.barrett_reduce((q - 1) ** 2, q, k, mu) == (q - 1) ** 2 % q


def test_trace_hook_matches_reference_log():
    g = load_golden("dilithium")
    q, omega = g["q"], g["omega"]
    a = g["cases"]["seed0_identity"]["a"]
    lines = ["CG NTT start", f"  omega_n={omega} modulus={q}", f"  input(first 16)={a[:16]}",
             f"  bitrev(first 16)={O.bit_reverse_list(a)[:16]}"]

    def hook(stage, k, omega_s, out):
        lines.append(f"  stage={stage} k={k} omega_s={omega_s}")
        lines.append(f"  stage_out(first 16)={out[:16]}")

    O.cg_ntt(a, omega, q, trace=hook)
    assert lines == g["verbose_log_seed0"]


@pytest.mark.parametrize("tag", TAGS)
def test_reference_cpp_build_agrees_with_oracle(tag, coracle):
    """oracle/_ref (the reference's own C++ sources) vs our restatement on random rows."""
    if not RefLib.available(tag):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    lib = RefLib(tag)
    rng = np.random.default_rng(11)
    rows = 4
    a = rng.integers(0, lib.q, size=(rows, lib.n), dtype=np.uint64)
    b = rng.integers(0, lib.q, size=(rows, lib.n), dtype=np.uint64)
    got = lib.polymul(a, b, threads=2).astype(np.uint64)
    want = coracle.nwc_poly_mult(a, b, lib.psi, lib.q, threads=2)
    assert (got == want).all()
    assert lib.checksum(lib.polymul(lib.make_poly(1), lib.make_poly(2))) == SURVEY_CHECKSUMS[tag][1]


def test_oracle_matches_the_reference_refs_twins():
    # test/refs/ntt_forward_reference.py / ntt_inverse_reference.py: cg_ntt over psi^2
    # with inputs reduced first; vectors generated by the reference (tests/golden/make_golden_refs.py)
    with open(os.path.join(GOLDEN_DIR, "golden_refs.json")) as fh:
        cases = json.load(fh)
    assert len(cases) >= 4
    for c in cases:
        n, q, omega = c["n"], c["q"], pow(c["psi"], 2, c["q"])
        x = [v % q for v in c["x"]]
        assert O.cg_ntt(x, omega, q) == c["forward"]
        assert O.cg_intt(x, omega, q) == c["inverse"]


def test_oracle_follows_the_reference_outside_the_ntt_friendly_domain():
    """Even / composite moduli, non-primitive roots, omega = 0: outputs of the reference's own cg_ntt.py
    (tests/golden/make_golden_domain.py) against the oracle's restatement."""
    import json
    import os

    from oracle import ntt_oracle as O

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_domain.json")
    with open(path) as fh:
        cases = json.load(fh)["cases"]
    assert len(cases) >= 8
    for c in cases:
        assert O.cg_ntt(c["a"], c["omega"], c["q"]) == c["cg_ntt"], c["note"]
        assert O.cg_intt(c["a"], c["omega"], c["q"]) == c["cg_intt"], c["note"]
        assert O.cg_intt(c["cg_ntt"], c["omega"], c["q"]) == c["roundtrip"], c["note"]
        assert O.nwc_poly_mult(c["a"], c["b"], c["psi"], c["q"]) == c["nwc_poly_mult"], c["note"]
