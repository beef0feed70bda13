#!/usr/bin/env python3
"""Multi-modulus throughput next to the single-modulus kernel (VERDICT r1 task 5): [L, B, N] in ONE launch per 16 limbs.

    python tools/rns_bench.py [--n 4096] [--limbs 1,4,16,32] [--rows 2048] [--bits 60]

Prints one JSON line per L: limb-polymul/s of tntt_rns_polymul over [L, rows, N] (CUDA events, inputs larger than L2
for L * rows >= 2048 at N = 4096), the same number of rows through the single-modulus plan of limb 0 (same kernel
shape family: the generic lazy reduction, not the Solinas special case), their ratio, the plan-creation time with
device-generated tables next to L single-modulus plans with host-built tables, and the multiplier-pipe fraction
from the survey's IMAD32 accounting.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tiny-ntt_b200"))
import torch  # noqa: E402

import tntt  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--limbs", default="1,4,16,32")
ap.add_argument("--rows", type=int, default=2048)
ap.add_argument("--bits", type=int, default=60)
ap.add_argument("--reps", type=int, default=20)
args = ap.parse_args()
n = args.n


def is_prime(m):
    if m < 2:
        return False
    for p in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if m % p == 0:
            return m == p
    d, r = m - 1, 0
    while d % 2 == 0:
        d, r = d // 2, r + 1
    for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        x = pow(a, d, m)
        if x in (1, m - 1):
            continue
        for _ in range(r - 1):
            x = x * x % m
            if x == m - 1:
                break
        else:
            return False
    return True


def primes(count, below):
    out, q = [], below - (below - 1) % (2 * n)
    while len(out) < count:
        if q < below and is_prime(q) and q != (1 << 60) - (1 << 14) + 1:     # skip the Solinas prime: generic kernels on both sides
            out.append(q)
        q -= 2 * n
    return out


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


MODMULS = {256: 3584, 1024: 17408, 4096: 81920}
peak_lo = tntt.microbench(0, 0)
for L in [int(v) for v in args.limbs.split(",")]:
    qs = primes(L, 1 << args.bits)
    t0 = time.perf_counter()
    psis = [tntt.find_psi(n, q) for q in qs]
    t_psi = time.perf_counter() - t0
    t0 = time.perf_counter()
    ctx = tntt.RnsContext(n, qs, psis)
    torch.cuda.synchronize()
    t_plan = time.perf_counter() - t0
    wb = ctx.word_bytes
    gen = torch.Generator(device="cuda").manual_seed(L)
    a = torch.stack([torch.randint(0, q, (args.rows, n), generator=gen, device="cuda", dtype=torch.int64) for q in qs]).to(ctx.dtype)
    b = torch.stack([torch.randint(0, q, (args.rows, n), generator=gen, device="cuda", dtype=torch.int64) for q in qs]).to(ctx.dtype)
    c = torch.empty_like(a)
    s_rns = timed(lambda: ctx.polymul(a, b, out=c), args.reps)
    t0 = time.perf_counter()
    plans = ctx.plans                     # L single-modulus plans, host-built tables
    torch.cuda.synchronize()
    t_single_plans = time.perf_counter() - t0
    fa, fb, fc = a.reshape(-1, n), b.reshape(-1, n), c.reshape(-1, n)
    s_single = timed(lambda: tntt.polymul(plans[0], fa, fb, out=fc), args.reps)       # same rows, one modulus, one launch
    s_loop = timed(lambda: [tntt.polymul(plans[l], a[l], b[l], out=c[l]) for l in range(L)], args.reps)
    total = L * args.rows
    imad = MODMULS[n] * (10.05 if wb == 8 else 3.0)
    print(json.dumps({
        "rns": True, "n": n, "limbs": L, "rows_per_limb": args.rows, "modulus_bits": args.bits, "word_bytes": wb,
        "kernel": ctx.kernel, "kernel_attributes": ctx.kernel_attributes(),
        "limb_polymul_per_s": total / s_rns, "single_modulus_same_rows_per_s": total / s_single,
        "ratio_vs_single_modulus": s_single / s_rns, "per_limb_launch_loop_per_s": total / s_loop,
        "launches_per_call": (L + 15) // 16, "imad32_frac": total / s_rns * imad / peak_lo,
        "plan_create_s_device_tables": t_plan, "find_psi_s": t_psi, "single_plans_create_s_host_tables": t_single_plans,
        "table_bytes": ctx.table_bytes,
    }), flush=True)
    del ctx, a, b, c, plans
    tntt.clear_plan_cache()
