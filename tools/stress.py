#!/usr/bin/env python3
"""Randomised parity stress: random NTT-friendly primes (24..60 bits), every fused size, random batch sizes,
every variant and the transform-domain / natural-order routes against the C oracle.  usage: stress.py [SECONDS]"""
import faulthandler
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tiny-ntt_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import tntt  # noqa: E402
from oracle.cpu_ref import COracle  # noqa: E402


def is_prime(n):
    if n < 2:
        return False
    for p in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if n % p == 0:
            return n == p
    d, r = n - 1, 0
    while d % 2 == 0:
        d //= 2
        r += 1
    for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(r - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    rnd = random.Random(2026)
    co = COracle()
    t0, cases, bigs, rns_cases, literal_cases = time.time(), 0, 0, 0, 0
    while time.time() - t0 < budget:
        # a case that takes longer than two minutes is a hang: dump every thread's stack and give up
        faulthandler.dump_traceback_later(120, exit=True)
        if cases % 500 == 0:
            print(f"[{time.time() - t0:6.1f} s] {cases} cases", file=sys.stderr, flush=True)
        logn = rnd.choice([8, 9, 10, 11, 12, 13, 14, 15])
        n = 1 << logn
        bits = rnd.choice([20, 23, 26, 28, 31, 40, 50, 58, 59, 60])
        if bits <= logn + 2:
            continue
        while True:
            k = rnd.randrange(1 << (bits - logn - 2), 1 << (bits - logn - 1))
            q = k * 2 * n + 1
            if q.bit_length() == bits and is_prime(q):
                break
        psi = tntt.find_psi(n, q)
        tntt.clear_plan_cache()
        plan = tntt.get_plan(n, q, psi, True)
        npdt, sdt = (np.uint32, np.int32) if plan.word_bytes == 4 else (np.uint64, np.int64)
        rows = rnd.choice([1, 2, 3, 17, 37, 38, 149, 200]) if logn <= 13 else rnd.choice([1, 3, 9])
        rng = np.random.default_rng(cases)
        a = rng.integers(0, q, size=(rows, n), dtype=np.uint64)
        b = rng.integers(0, q, size=(rows, n), dtype=np.uint64)
        a[0] = q - 1
        if rows > 1:
            b[1] = q - 1
        ta = torch.from_numpy(a.astype(npdt).view(sdt)).cuda()
        tb = torch.from_numpy(b.astype(npdt).view(sdt)).cuda()
        host = lambda t: t.cpu().numpy().view(npdt).astype(np.uint64)      # noqa: E731
        want = co.nwc_poly_mult(a, b, psi, q, threads=8)
        assert (host(tntt.polymul(plan, ta, tb)) == want).all(), ("dispatch", n, q, rows)
        for v, d in plan.variants():
            assert (host(tntt.polymul(plan, ta, tb, variant=v)) == want).all(), (d, n, q, rows)
        if plan.spectrum:
            sb = tntt.forward_spectrum(plan, tb)
            assert (host(tntt.polymul_spectrum(plan, ta, sb)) == want).all(), ("spectrum", n, q, rows)
            assert (host(tntt.inverse_spectrum(plan, tntt.pointwise(plan, tntt.forward_spectrum(plan, ta), sb))) == want).all()
        omega = psi * psi % q
        assert (host(tntt.forward(plan, ta)) == co.cg_ntt(a, omega, q)).all(), ("cg_ntt", n, q, rows)
        assert (host(tntt.inverse(plan, ta)) == co.cg_intt(a, omega, q)).all(), ("cg_intt", n, q, rows)
        # twisted natural-order transforms (the permuted tile paths) against the fused product
        if plan.spectrum:
            fa, fb = tntt.forward(plan, ta, twist=True), tntt.forward(plan, tb, twist=True)
            assert (host(tntt.inverse(plan, tntt.pointwise(plan, fa, fb), twist=True)) == want).all(), ("twisted", n, q, rows)
        # every few cases: a batch that fills the GPU several times over (ordering bugs between warps / CTAs only show
        # under load), checked by identities between independent kernel paths instead of the CPU oracle
        if cases % 4 == 0 and logn <= 13:
            big = max(2048, (1 << 24) >> logn)
            g = torch.Generator(device="cuda").manual_seed(cases)
            xa = torch.randint(0, q, (big, n), generator=g, device="cuda", dtype=torch.int64).to(plan.dtype)
            xb = torch.randint(0, q, (big, n), generator=g, device="cuda", dtype=torch.int64).to(plan.dtype)
            c0 = tntt.polymul(plan, xa, xb)
            for v, d in plan.variants():
                assert torch.equal(tntt.polymul(plan, xa, xb, variant=v), c0), ("big batch, variant", d, n, q)
            assert torch.equal(tntt.inverse(plan, tntt.forward(plan, xa)), xa), ("big batch, round trip", n, q)
            if plan.spectrum:
                sb = tntt.forward_spectrum(plan, xb)
                assert torch.equal(tntt.polymul_spectrum(plan, xa, sb), c0), ("big batch, spectrum", n, q)
                assert torch.equal(tntt.inverse_spectrum(plan, sb), xb), ("big batch, spectrum round trip", n, q)
                fa, fb = tntt.forward(plan, xa, twist=True), tntt.forward(plan, xb, twist=True)
                assert torch.equal(tntt.inverse(plan, tntt.pointwise(plan, fa, fb), twist=True), c0), ("big batch, twisted", n, q)
            del xa, xb, c0
            bigs += 1
        # every few cases: a multi-modulus context over random primes of one word class, against the oracle limb by limb
        if cases % 7 == 0 and logn in (8, 10, 12):
            L = rnd.choice([1, 2, 5, 16, 17, 20])
            qs, k0 = [], q
            while len(qs) < L and k0 > 2 * n:
                if is_prime(k0):
                    qs.append(k0)
                k0 -= 2 * n
            L = len(qs)
            try:
                ctx = tntt.RnsContext(n, qs)
            except ValueError:
                ctx = None          # the descending primes straddle a word / reduction class boundary: not a valid context
            if ctx is not None:
                r2 = rnd.choice([1, 3, 33])
                ra = np.stack([rng.integers(0, m, size=(r2, n), dtype=np.uint64) for m in qs])
                rb = np.stack([rng.integers(0, m, size=(r2, n), dtype=np.uint64) for m in qs])
                nd, sd = (np.uint32, np.int32) if ctx.word_bytes == 4 else (np.uint64, np.int64)
                da = torch.from_numpy(ra.astype(nd).view(sd)).cuda()
                db = torch.from_numpy(rb.astype(nd).view(sd)).cuda()
                got = ctx.polymul(da, db)
                assert torch.equal(ctx.polymul_spectrum(da, ctx.forward_spectrum(db)), got), ("rns spectrum", n, qs[0], L)
                gh = got.cpu().numpy().view(nd).astype(np.uint64)
                for l, (m, ps) in enumerate(zip(ctx.moduli, ctx.psis)):
                    assert (gh[l] == co.nwc_poly_mult(ra[l], rb[l], ps, m, threads=8)).all(), ("rns", n, m, L, l)
                    assert ctx.tables_match_host_generators(l)
                rns_cases += 1
                del ctx
        # ... and a literal plan: a modulus that is no prime (the reference accepts it), small n, against the Python oracle
        if cases % 11 == 0:
            from oracle import ntt_oracle as PO
            ln = rnd.choice([3, 4, 5, 6])
            m = rnd.randrange(2, 1 << rnd.choice([8, 20, 31, 45, 59])) | rnd.choice([0, 1])
            w = rnd.randrange(m)
            lp = tntt.get_plan(1 << ln, m, w, True)
            xs = [rnd.randrange(m) for _ in range(1 << ln)]
            ys = [rnd.randrange(m) for _ in range(1 << ln)]
            nd, sd = (np.uint32, np.int32) if lp.word_bytes == 4 else (np.uint64, np.int64)
            tx = torch.from_numpy(np.array([xs], dtype=nd).view(sd)).cuda()
            ty = torch.from_numpy(np.array([ys], dtype=nd).view(sd)).cuda()
            got = tntt.polymul(lp, tx, ty).cpu().numpy().view(nd).astype(np.uint64)[0].tolist()
            assert got == PO.nwc_poly_mult(xs, ys, w, m), ("literal", 1 << ln, m, w)
            literal_cases += 1
        cases += 1
    faulthandler.cancel_dump_traceback_later()
    print(f"stress ok: {cases} random (n, q, batch) cases ({bigs} with GPU-filling batches, {rns_cases} multi-modulus contexts, "
          f"{literal_cases} literal plans) in {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
