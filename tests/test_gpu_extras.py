"""GPU tests of the rows next to the hot path: the plain-C client, RNS batches, the TMA variant."""
import os
import subprocess

import numpy as np
import pytest

from oracle import ntt_oracle as O
from oracle.cpu_ref import COracle

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plain_c_client_follows_the_rocc_sequence(tmp_path):
    """examples/rocc_style_driver.c: load A, load B, start, read -- the vector of chipyard/ntt-test.c:101-107."""
    exe = str(tmp_path / "rocc_driver")
    lib = os.path.join(ROOT, "tiny-ntt_b200")
    subprocess.check_call(["gcc", "-O2", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "rocc_style_driver.c"),
                           "-L" + lib, "-ltntt", "-Wl,-rpath," + lib, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "c[0..3] = 5 11 17 3" in out.stdout and "PASS" in out.stdout


def test_rns_product_matches_bigint_schoolbook():
    import tntt

    n = 256
    # three NTT-friendly 60-bit primes q = k * 2^20 + 1 below 2^60 (found by trial, checked prime by the library)
    moduli = []
    k = (1 << 40) - 1
    while len(moduli) < 3:
        q = k * (1 << 20) + 1
        if q < (1 << 60) and all(q % p for p in (3, 5, 7, 11, 13)) and pow(2, q - 1, q) == 1:
            try:
                tntt.find_psi(n, q)
                moduli.append(q)
            except ValueError:
                pass
        k -= 1
    ctx = tntt.RnsContext(n, moduli)
    assert ctx.word_bytes == 8
    rng = np.random.default_rng(3)
    rows = 3
    big_a = [[int(rng.integers(0, 1 << 62)) * int(rng.integers(0, 1 << 62)) % ctx.Q for _ in range(n)] for _ in range(rows)]
    big_b = [[int(rng.integers(0, 1 << 62)) for _ in range(n)] for _ in range(rows)]
    a = torch.from_numpy(ctx.decompose(big_a).view(np.int64)).cuda()
    b = torch.from_numpy(ctx.decompose(big_b).view(np.int64)).cuda()
    c = ctx.polymul(a, b)
    got = ctx.reconstruct(c.cpu().numpy().view(np.uint64))
    for r in range(rows):
        assert got[r] == O.schoolbook_negacyclic(big_a[r], big_b[r], ctx.Q)
    # transform-domain round trip per limb
    assert torch.equal(ctx.inverse(ctx.forward(a)), a)


def test_tma_variant_agrees_and_is_listed():
    import tntt

    p = O.PARAMS["n4096_60"]
    plan = tntt.get_plan(p["n"], p["q"], p["psi"], True)
    tma = [v for v, d in plan.variants() if d.split(" ")[0].endswith("_t1")]
    assert tma, "no TMA-staged variant built"
    rng = np.random.default_rng(8)
    a = rng.integers(0, p["q"], size=(300, p["n"]), dtype=np.uint64)
    b = rng.integers(0, p["q"], size=(300, p["n"]), dtype=np.uint64)
    want = COracle().nwc_poly_mult(a, b, p["psi"], p["q"], threads=8)
    da, db = torch.from_numpy(a.view(np.int64)).cuda(), torch.from_numpy(b.view(np.int64)).cuda()
    for v in tma:
        for _ in range(3):   # back-to-back launches reuse the mbarrier phase logic from scratch each time
            got = tntt.polymul(plan, da, db, variant=v).cpu().numpy().view(np.uint64)
            assert (got == want).all()
