"""The kernels' index maps, tables and lazy arithmetic, executed on the CPU (tests/host_emul.cpp
replays the kernel bodies from the same __host__ __device__ code) and compared with the oracle."""
import random

import numpy as np
import pytest

import emu
from oracle import ntt_oracle as O
from oracle.cpu_ref import COracle

#        word logn logr ppc na red tag
POLY = [(4, 8, 4, 16, 1, 0, "dilithium"), (4, 8, 4, 16, 2, 0, "dilithium"), (4, 8, 3, 8, 1, 0, "dilithium"),
        (4, 10, 5, 8, 1, 0, "n1024_24"), (4, 10, 5, 8, 2, 0, "n1024_24"), (4, 10, 4, 4, 1, 0, "n1024_24"),
        (4, 10, 4, 4, 2, 0, "n1024_24"), (4, 12, 4, 1, 1, 0, "n4096_24"), (4, 12, 4, 1, 2, 0, "n4096_24"),
        (4, 12, 5, 2, 1, 0, "n4096_24"), (4, 12, 3, 1, 1, 0, "n4096_24"),
        (8, 12, 4, 1, 1, 1, "n4096_60"), (8, 12, 4, 1, 2, 1, "n4096_60"), (8, 12, 3, 1, 1, 1, "n4096_60"),
        (8, 12, 3, 1, 2, 1, "n4096_60"), (8, 12, 4, 1, 1, 0, "n4096_24"), (8, 8, 4, 16, 1, 0, "dilithium"),
        (8, 8, 4, 16, 1, 1, "dilithium"), (8, 10, 4, 4, 1, 0, "n1024_24"), (8, 10, 4, 4, 1, 1, "n1024_24")]


@pytest.fixture(scope="module")
def co():
    return COracle()


@pytest.mark.parametrize("wb,logn,logr,ppc,na,red,tag", POLY)
def test_emulated_polymul_matches_oracle(wb, logn, logr, ppc, na, red, tag, co):
    p = O.PARAMS[tag]
    n, q, psi = p["n"], p["q"], p["psi"]
    rng = np.random.default_rng(logn * 100 + logr * 10 + na)
    batch = ppc + 3  # ragged: the last CTA is partly empty
    a = rng.integers(0, q, size=(batch, n), dtype=np.uint64)
    b = rng.integers(0, q, size=(batch, n), dtype=np.uint64)
    a[0], b[0] = O.make_poly(tag, 1), O.make_poly(tag, 2)
    a[1], b[1] = q - 1, q - 1          # largest canonical values: worst case for the lazy ranges
    a[2], b[2] = 0, q - 1
    want = co.nwc_poly_mult(a, b, psi, q, threads=4)
    got = emu.polymul(wb, logn, logr, ppc, na, red, a, b, q, psi).astype(np.uint64)
    assert (got == want).all()
    assert emu.lib().emu_range_violations() == 0     # no lazy value ever wrapped around the word


def test_emulated_tiny_sizes(co):
    # n = 4 worked example (test/refs/fast_ntt_negacyclic_convolution.py:161-214) and n = 16, 32
    got = emu.polymul(4, 2, 1, 2, 1, 0, [[1, 2, 3, 4]], [[5, 6, 7, 8]], 7681, 1925)
    assert got.tolist() == [[7625, 7645, 2, 60]]
    q = 8380417
    for logn, logr, ppc, na in ((4, 2, 4, 1), (5, 2, 2, 2)):
        n = 1 << logn
        psi = pow(1239911, 256 // n, q)
        rng = np.random.default_rng(n)
        a = rng.integers(0, q, size=(5, n), dtype=np.uint64)
        b = rng.integers(0, q, size=(5, n), dtype=np.uint64)
        got = emu.polymul(4, logn, logr, ppc, na, 0, a, b, q, psi).astype(np.uint64)
        assert (got == co.nwc_poly_mult(a, b, psi, q)).all()


XF = [(4, 8, 4, 16, 0, "dilithium"), (4, 10, 5, 8, 0, "n1024_24"), (4, 10, 4, 4, 0, "n1024_24"),
      (4, 12, 4, 1, 0, "n4096_24"), (8, 8, 4, 16, 1, "dilithium"), (8, 12, 4, 1, 0, "n4096_24"),
      (8, 12, 4, 1, 1, "n4096_60"), (8, 12, 3, 1, 1, "n4096_60")]


@pytest.mark.parametrize("wb,logn,logr,ppc,red,tag", XF)
def test_emulated_transforms_match_oracle(wb, logn, logr, ppc, red, tag, co):
    p = O.PARAMS[tag]
    n, q, psi = p["n"], p["q"], p["psi"]
    omega = psi * psi % q
    rng = np.random.default_rng(5)
    x = rng.integers(0, q, size=(ppc + 1, n), dtype=np.uint64)
    x[0] = q - 1
    fwd = emu.transform(wb, logn, logr, ppc, red, x, q, omega, 0).astype(np.uint64)
    assert (fwd == co.cg_ntt(x, omega, q)).all()                         # cg_ntt
    inv = emu.transform(wb, logn, logr, ppc, red, x, q, omega, 1).astype(np.uint64)
    assert (inv == co.cg_intt(x, omega, q)).all()                        # cg_intt
    back = emu.transform(wb, logn, logr, ppc, red, fwd, q, omega, 1).astype(np.uint64)
    assert (back == x).all()                                             # round trip
    tw = emu.transform(wb, logn, logr, ppc, red, x[:1], q, psi, 2).astype(np.uint64)
    assert tw[0].tolist() == O.forward_negacyclic([int(v) for v in x[0]], psi, q)   # ntt(twist(a))
    rt = emu.transform(wb, logn, logr, ppc, red, tw, q, psi, 3).astype(np.uint64)
    assert (rt == x[:1]).all()                                           # untwist(intt(...))
    assert emu.lib().emu_range_violations() == 0


def test_emulated_reduce_input_flag():
    q = 8380417
    omega = pow(1239911, 2, q)
    rng = np.random.default_rng(9)
    raw = rng.integers(0, 2**32, size=(2, 256), dtype=np.uint64)
    got = emu.transform(4, 8, 4, 16, 0, raw, q, omega, 0, reduce_input=1)
    want = [O.cg_ntt([int(v) for v in row], omega, q) for row in raw]
    assert got.tolist() == want


def test_modular_arithmetic_primitives():
    L = emu.lib()
    rnd = random.Random(3)
    q60, q24 = O.PARAMS["n4096_60"]["q"], 8380417
    r64, r32 = pow(2, 64, q60), pow(2, 32, q24)
    for _ in range(20000):
        x, w = rnd.getrandbits(64), rnd.randrange(q60)
        t = L.emu_shoup64(x, w, q60)
        assert t < 2 * q60 and t % q60 == x * w % q60            # any 64-bit x -> [0, 2q)
        t = L.emu_shoup_lazy64(x, w, q60)
        assert t < 3 * q60 and t % q60 == x * w % q60            # butterfly product: [0, 3q)
        a, b = rnd.randrange(1 << 63), rnd.randrange(1 << 63)
        m = L.emu_mont64(a, b, q60)
        assert m < (1 << 62) + q60 and m * r64 % q60 == a * b % q60
        c, d = rnd.randrange(q60), rnd.randrange(q60)
        assert L.emu_barrett64(c, d, q60) == c * d % q60         # rtl/barrett_reduction.v formula
        v = rnd.getrandbits(64)
        s = L.emu_csub_top64(v, q60)
        assert s % q60 == v % q60 and s < (1 << 63) + q60 * 8
        x, w = rnd.getrandbits(32), rnd.randrange(q24)
        t = L.emu_shoup32(x, w, q24)
        assert t < 2 * q24 and t % q24 == x * w % q24
        a, b = rnd.randrange(25 * q24), rnd.randrange(25 * q24)
        m = L.emu_mont32(a, b, q24)
        assert m * r32 % q24 == a * b % q24 and m < 3 * q24
        c, d = rnd.randrange(q24), rnd.randrange(q24)
        assert L.emu_barrett32(c, d, q24) == c * d % q24
    assert L.emu_barrett64(q60 - 1, q60 - 1, q60) == (q60 - 1) ** 2 % q60
    assert L.emu_barrett32(q24 - 1, q24 - 1, q24) == (q24 - 1) ** 2 % q24
    # other moduli exercise the general shifts
    for q in (7681, 12289, 65537, 2013265921, 4611686018326724609 >> 3 | 1):
        k = q.bit_length()
        for _ in range(2000):
            c, d = rnd.randrange(q), rnd.randrange(q)
            assert L.emu_barrett64(c, d, q) == c * d % q


def test_host_number_theory_helpers():
    L = emu.lib()
    for p in (2, 3, 7681, 8380417, 1152921504606830593, 2013265921):
        assert L.emu_is_prime(p)
    for c in (1, 9, 8380417 * 3, 1152921504606830593 - 2, 2**59):
        assert not L.emu_is_prime(c)
    assert L.emu_lazy_full_ok(4, 8380417, 12) == 1            # Dilithium modulus runs reduction-free in 32 bits
    assert L.emu_lazy_full_ok(4, (1 << 30) - 35, 12) == 0
    assert L.emu_lazy_full_ok(8, 1152921504606830593, 12) == 0  # the 60-bit modulus needs the per-pass reduction
    assert L.emu_lazy_full_ok(8, (1 << 50) - 27, 12) == 1


SPEC = [(4, 8, 4, 16, 0, "dilithium"), (4, 10, 5, 8, 0, "n1024_24"), (4, 12, 4, 1, 0, "n4096_24"),
        (8, 8, 4, 16, 1, "dilithium"), (8, 10, 4, 4, 0, "n1024_24"), (8, 12, 4, 1, 0, "n4096_24"),
        (8, 12, 4, 1, 1, "n4096_60")]


@pytest.mark.parametrize("wb,logn,logr,ppc,red,tag", SPEC)
def test_emulated_transform_domain_kernels(wb, logn, logr, ppc, red, tag, co):
    # spectrum_forward / spectrum_inverse / polymul_spectrum kernel bodies (csrc/kernels.cuh) replayed on the CPU
    p = O.PARAMS[tag]
    n, q, psi = p["n"], p["q"], p["psi"]
    rng = np.random.default_rng(logn * 7 + wb)
    batch = ppc + 2
    a = rng.integers(0, q, size=(batch, n), dtype=np.uint64)
    b = rng.integers(0, q, size=(batch, n), dtype=np.uint64)
    a[0], b[0] = q - 1, q - 1
    a[1], b[1] = O.make_poly(tag, 1), O.make_poly(tag, 2)
    assert (emu.spectrum(wb, logn, logr, ppc, red, a, b, q, psi, 0).astype(np.uint64) == a).all()          # round trip
    want = co.nwc_poly_mult(a, b, psi, q, threads=4)
    assert (emu.spectrum(wb, logn, logr, ppc, red, a, b, q, psi, 1).astype(np.uint64) == want).all()       # per-row spectra
    shared = co.nwc_poly_mult(a, np.broadcast_to(b[0], a.shape).copy(), psi, q, threads=4)
    assert (emu.spectrum(wb, logn, logr, ppc, red, a, b, q, psi, 2).astype(np.uint64) == shared).all()     # one shared spectrum
    spec = emu.spectrum(wb, logn, logr, ppc, red, a[:1], b[:1], q, psi, 3).astype(np.uint64)[0]
    assert spec.max() < q                                                                                   # canonical
    assert sorted(spec.tolist()) == sorted(O.forward_negacyclic([int(v) for v in a[0]], psi, q))            # a permutation of ntt(twist(a))
    assert emu.lib().emu_range_violations() == 0


@pytest.mark.parametrize("wb,logn,logr,ppc,red,tag", SPEC)
def test_emulated_natural_order_transforms_on_the_fused_passes(wb, logn, logr, ppc, red, tag, co):
    # tntt_forward / tntt_inverse for the fused sizes: Cooley-Tukey passes + permuted store, permuted load + DIT
    p = O.PARAMS[tag]
    n, q, psi = p["n"], p["q"], p["psi"]
    omega = psi * psi % q
    rng = np.random.default_rng(logn + wb)
    x = rng.integers(0, q, size=(ppc + 1, n), dtype=np.uint64)
    x[0] = q - 1
    fwd = emu.spectrum(wb, logn, logr, ppc, red, x, x, q, psi, 5).astype(np.uint64)
    assert (fwd == co.cg_ntt(x, omega, q)).all()                                               # cg_ntt
    assert (emu.spectrum(wb, logn, logr, ppc, red, x, x, q, psi, 7).astype(np.uint64) == co.cg_intt(x, omega, q)).all()   # cg_intt
    assert (emu.spectrum(wb, logn, logr, ppc, red, fwd, fwd, q, psi, 7).astype(np.uint64) == x).all()
    tw = emu.spectrum(wb, logn, logr, ppc, red, x[:1], x[:1], q, psi, 4).astype(np.uint64)
    assert tw[0].tolist() == O.forward_negacyclic([int(v) for v in x[0]], psi, q)              # ntt(twist(a)), natural order
    assert (emu.spectrum(wb, logn, logr, ppc, red, tw, tw, q, psi, 6).astype(np.uint64) == x[:1]).all()
    assert emu.lib().emu_range_violations() == 0


# sizes next to the reference's three (other NTT-friendly rings): (word, logn, logr, ppc, na, red, q, psi)
Q60 = (1 << 60) - (1 << 14) + 1
EXTRA_SIZES = [
    (4, 9, 5, 16, 2, 0, 8380417, 1718063), (4, 11, 4, 2, 2, 0, 8380417, 7901702), (4, 13, 5, 1, 2, 0, 67043329, 8157893),
    (8, 9, 4, 8, 1, 0, 8380417, 1718063), (8, 9, 4, 8, 1, 1, Q60, 984081769261068913),
    (8, 11, 4, 2, 1, 0, 8380417, 7901702), (8, 11, 4, 2, 1, 1, Q60, 644283108363935541),
    (8, 11, 4, 2, 2, 1, Q60, 644283108363935541),
    (8, 13, 4, 1, 1, 0, 67043329, 8157893), (8, 13, 4, 1, 1, 1, Q60, 527760526715669589),
]


@pytest.mark.parametrize("wb,logn,logr,ppc,na,red,q,psi", EXTRA_SIZES)
def test_emulated_kernels_at_other_sizes(wb, logn, logr, ppc, na, red, q, psi, co):
    n = 1 << logn
    assert pow(psi, n, q) == q - 1
    omega = psi * psi % q
    rng = np.random.default_rng(logn)
    batch = ppc + 1
    a = rng.integers(0, q, size=(batch, n), dtype=np.uint64)
    b = rng.integers(0, q, size=(batch, n), dtype=np.uint64)
    a[0], b[0] = q - 1, q - 1
    want = co.nwc_poly_mult(a, b, psi, q, threads=4)
    assert (emu.polymul(wb, logn, logr, ppc, na, red, a, b, q, psi).astype(np.uint64) == want).all()
    assert (emu.spectrum(wb, logn, logr, ppc, red, a, b, q, psi, 1).astype(np.uint64) == want).all()
    assert (emu.spectrum(wb, logn, logr, ppc, red, a, b, q, psi, 0).astype(np.uint64) == a).all()
    assert (emu.spectrum(wb, logn, logr, ppc, red, a, a, q, psi, 5).astype(np.uint64) == co.cg_ntt(a, omega, q)).all()
    assert (emu.spectrum(wb, logn, logr, ppc, red, a, a, q, psi, 7).astype(np.uint64) == co.cg_intt(a, omega, q)).all()
    assert emu.lib().emu_range_violations() == 0


_largest_friendly_prime = emu.largest_friendly_prime


#                 word logn logr ppc na red
BOUNDARY_SHAPES = [(4, 8, 4, 16, 2, 0), (4, 10, 5, 8, 2, 0), (4, 12, 4, 1, 2, 0), (4, 13, 5, 1, 2, 0),
                   (8, 8, 4, 16, 1, 0), (8, 12, 4, 1, 1, 0), (8, 13, 4, 1, 1, 0),
                   (8, 8, 4, 16, 1, 1), (8, 12, 4, 1, 1, 1), (8, 12, 4, 1, 2, 1), (8, 13, 4, 1, 1, 1)]


@pytest.mark.parametrize("wb,logn,logr,ppc,na,red", BOUNDARY_SHAPES)
def test_emulated_kernels_at_the_largest_modulus_each_path_accepts(wb, logn, logr, ppc, na, red, co):
    # the lazy value ranges are tightest for the largest modulus a path admits: the largest NTT-friendly prime that
    # still passes lazy_full_ok (reduction-free paths) or is below 2^60 (paths with per-pass reduction), with
    # all-(q-1) rows; the emulation audits every butterfly for wrap-around
    from tntt.rns import find_psi

    L = emu.lib()
    n = 1 << logn
    ok = (lambda q: q < (1 << 60)) if red else (lambda q: bool(L.emu_lazy_full_ok(wb, q, logn)))
    q = _largest_friendly_prime(n, ok)
    assert ok(q) and (red or not ok(q + 2 * n * 64) or q > (1 << 59))
    psi = find_psi(n, q)
    rng = np.random.default_rng(logn + wb + red)
    batch = ppc + 1
    a = rng.integers(0, q, size=(batch, n), dtype=np.uint64)
    b = rng.integers(0, q, size=(batch, n), dtype=np.uint64)
    a[0], b[0] = q - 1, q - 1
    a[1] = q - 1
    want = co.nwc_poly_mult(a, b, psi, q, threads=4)
    assert (emu.polymul(wb, logn, logr, ppc, na, red, a, b, q, psi).astype(np.uint64) == want).all(), q
    assert emu.lib().emu_range_violations() == 0, q
    if na == 1 or (wb, logn) in ((4, 8), (4, 10), (4, 12), (4, 13)):     # the plan's transform-domain shape
        assert (emu.spectrum(wb, logn, logr, ppc, red, a, b, q, psi, 1).astype(np.uint64) == want).all(), q
        assert (emu.spectrum(wb, logn, logr, ppc, red, a, b, q, psi, 0).astype(np.uint64) == a).all(), q
        assert (emu.spectrum(wb, logn, logr, ppc, red, a, a, q, psi, 7).astype(np.uint64) == co.cg_intt(a, psi * psi % q, q)).all(), q
        assert emu.lib().emu_range_violations() == 0, q


# round 2: padded tile (pad) and the Solinas-form reductions (red 2) of the N = 4096 / 60-bit shapes
SOLINAS_Q = (1 << 60) - (1 << 14) + 1


@pytest.mark.parametrize("na,red,pad", [(2, 1, 1), (2, 2, 1), (1, 2, 0), (2, 2, 0), (1, 1, 1), (1, 2, 1)])
def test_emulated_padded_and_solinas_shapes(na, red, pad, co):
    p = O.PARAMS["n4096_60"]
    n, q, psi = p["n"], p["q"], p["psi"]
    assert q == SOLINAS_Q
    rng = np.random.default_rng(na * 100 + red * 10 + pad)
    a = rng.integers(0, q, size=(4, n), dtype=np.uint64)
    b = rng.integers(0, q, size=(4, n), dtype=np.uint64)
    a[0], b[0] = O.make_poly("n4096_60", 1), O.make_poly("n4096_60", 2)
    a[1], b[1] = q - 1, q - 1          # largest canonical values: worst case for the lazy ranges
    a[2], b[2] = 0, q - 1
    want = co.nwc_poly_mult(a, b, psi, q, threads=4)
    got = emu.polymul(8, 12, 4, 1, na, red, a, b, q, psi, pad=pad).astype(np.uint64)
    assert (got == want).all()
    assert emu.lib().emu_range_violations() == 0


def test_solinas_shapes_refuse_other_moduli():
    # red 2 is written for one modulus; the emulation (like tntt_variant_matches) refuses every other
    q = emu.largest_friendly_prime(4096, lambda v: v < SOLINAS_Q)
    assert q != SOLINAS_Q and q % 8192 == 1
    a = np.zeros((1, 4096), dtype=np.uint64)
    with pytest.raises(RuntimeError):
        emu.polymul(8, 12, 4, 1, 2, 2, a, a, q, 3, pad=1)


def test_solinas_arithmetic_primitives():
    L = emu.lib()
    q = SOLINAS_Q
    rnd = random.Random(5)
    edge = [0, 1, q - 1, q, q + 1, 2 * q, (1 << 60) - 1, 1 << 60, (1 << 64) - 1, (1 << 64) - q, 15 * q + 12345]
    for x in edge + [rnd.getrandbits(64) for _ in range(20000)]:
        r = L.emu_solinas_reduce(x)
        assert r % q == x % q and r < q + (1 << 18)          # "below two units" of the bound tracker
    for _ in range(20000):
        u, v = rnd.getrandbits(64), rnd.getrandbits(64)
        r = L.emu_solinas_mul(u, v)
        assert r % q == u * v % q and r < (1 << 60) + (1 << 37)
    for u in edge:
        for v in edge:
            r = L.emu_solinas_mul(u, v)
            assert r % q == u * v % q and r < (1 << 60) + (1 << 37)


@pytest.mark.parametrize("logn,logr,ppc,na,pad,tag", [(12, 4, 1, 1, 0, "n4096_60"), (12, 4, 1, 2, 1, "n4096_60"), (8, 4, 16, 1, 0, "dilithium")])
def test_emulated_barrett_shapes(logn, logr, ppc, na, pad, tag, co):
    """red 3: the reference's arithmetic (rtl/barrett_reduction.v:23-29 products, fully reducing adds) inside the fused kernel"""
    p = O.PARAMS[tag]
    n, q, psi = p["n"], p["q"], p["psi"]
    rng = np.random.default_rng(logn + na)
    a = rng.integers(0, q, size=(ppc + 2, n), dtype=np.uint64)
    b = rng.integers(0, q, size=(ppc + 2, n), dtype=np.uint64)
    a[0], b[0] = q - 1, q - 1
    want = co.nwc_poly_mult(a, b, psi, q, threads=4)
    assert (emu.polymul(8, logn, logr, ppc, na, 3, a, b, q, psi, pad=pad).astype(np.uint64) == want).all()


def test_emulated_solinas_transform_domain_kernels(co):
    """sp_u64_n12_r4_p1_red2: the transform-domain kernels with the Solinas reductions (the 60-bit prime only)"""
    p = O.PARAMS["n4096_60"]
    n, q, psi = p["n"], p["q"], p["psi"]
    rng = np.random.default_rng(77)
    a = rng.integers(0, q, size=(3, n), dtype=np.uint64)
    b = rng.integers(0, q, size=(3, n), dtype=np.uint64)
    a[0], b[0] = q - 1, q - 1
    want = co.nwc_poly_mult(a, b, psi, q, threads=4)
    args = (8, 12, 4, 1, 2, a, b, q, psi)
    assert (emu.spectrum(*args, 0).astype(np.uint64) == a).all()
    assert (emu.spectrum(*args, 1).astype(np.uint64) == want).all()
    shared = co.nwc_poly_mult(a, np.broadcast_to(b[0], b.shape).copy(), psi, q, threads=4)
    assert (emu.spectrum(*args, 2).astype(np.uint64) == shared).all()
    # the spectrum itself is the same canonical row whichever reduction mode produced it
    assert (emu.spectrum(*args, 3) == emu.spectrum(8, 12, 4, 1, 1, a, b, q, psi, 3)).all()
    omega = psi * psi % q
    assert (emu.spectrum(*args, 5).astype(np.uint64) == co.cg_ntt(a, omega, q)).all()
    assert (emu.spectrum(*args, 7).astype(np.uint64) == co.cg_intt(a, omega, q)).all()
    assert emu.lib().emu_range_violations() == 0


@pytest.mark.parametrize("logn,logr,ppc,tag", [(8, 4, 16, "dilithium"), (10, 5, 8, "n1024_24"), (12, 4, 1, "n4096_24")])
def test_emulated_padded_32bit_shapes(logn, logr, ppc, tag, co):
    p = O.PARAMS[tag]
    n, q, psi = p["n"], p["q"], p["psi"]
    rng = np.random.default_rng(logn)
    a = rng.integers(0, q, size=(ppc + 3, n), dtype=np.uint64)
    b = rng.integers(0, q, size=(ppc + 3, n), dtype=np.uint64)
    a[0], b[0] = q - 1, q - 1
    want = co.nwc_poly_mult(a, b, psi, q, threads=4)
    assert (emu.polymul(4, logn, logr, ppc, 2, 0, a, b, q, psi, pad=1).astype(np.uint64) == want).all()
    assert emu.lib().emu_range_violations() == 0


def test_solinas_first_inverse_pass_bound_tracker_is_sound():
    """dit2_pass0_bounds / dit2_step (csrc/modarith.cuh): the j = 0 butterflies of the first inverse pass skip their product
    and DOUBLE a lazy bound.  Replay the tracker's own decisions on worst-case numbers -- every register at the top of its
    tracked range, products at the top of theirs -- and check that nothing reaches 2^64, that the constant added by a
    product-free butterfly covers its subtrahend, and that the tracked bound really bounds every value."""
    import ctypes as C
    L = emu.lib()
    q, G = SOLINAS_Q, 3
    max_stage = L.emu_dit2_j0_max_stage()
    assert max_stage >= 1                      # the shipped setting: stages 0 and 1 (profiles/r02_whatif_j0.log)
    for logr in (3, 4, 5):
        R = 1 << logr
        for b0 in (1, 2):
            vals = [b0 * q - 1] * R            # largest values the tracker allows at the input (strictly below b0 q)
            for B in range(logr):
                bounds = (C.c_int * 32)()
                L.emu_dit2_pass0_bounds(G, b0, logr, B, bounds)
                assert all(vals[k] < bounds[k] * q for k in range(R)), (logr, b0, B)
                for k0 in range(R):
                    if k0 & (1 << B):
                        continue
                    k1 = k0 | (1 << B)
                    trivial = (k0 & ((1 << B) - 1)) == 0 and B <= max_stage
                    st = (C.c_int * 4)()
                    L.emu_dit2_step(int(trivial), G, bounds[k0], bounds[k1], st)
                    bx, by, red_x, red_y = list(st)
                    x, y = vals[k0], vals[k1]
                    if red_x:
                        x = L.emu_solinas_reduce(x)
                        assert x < 2 * q and bx == 2
                    if red_y:
                        y = L.emu_solinas_reduce(y)
                        assert y < 2 * q and by == 2
                    if trivial:
                        c = by * q
                        assert y <= c and x + y < 1 << 64 and x + c < 1 << 64
                        # worst cases of the two outputs: x + y, and x - y + c with the smallest y
                        vals[k0], vals[k1] = x + y, x + c
                    else:
                        # a lazy product is below G q for ANY word y; outputs x + v and x - v + G q
                        assert x + G * q < 1 << 64
                        vals[k0], vals[k1] = x + G * q - 1, x + G * q      # v = G q - 1 and v = 0
            after = (C.c_int * 32)()
            L.emu_dit2_pass0_bounds(G, b0, logr, logr, after)
            assert all(vals[k] < after[k] * q for k in range(R)) and max(after[:R]) <= 16
            assert L.emu_dit2_bound_at(G, b0, logr, logr) == max(after[:R])
            # later stages continue from that bound with the uniform rule and never pass 16 units either
            assert all(L.emu_dit2_bound_at(G, b0, logr, s) <= 16 for s in range(logr, 16))
