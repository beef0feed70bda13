#!/usr/bin/env python3
"""Every fused variant at N = 16384 / 32768 (cluster kernels, 60-bit and 27-bit moduli), polymul/s."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tiny-ntt_b200"))
import torch  # noqa: E402

import tntt  # noqa: E402

RINGS = [(16384, 1152921504606748673, 641000223749548346), (32768, 1152921504606584833, 1100972123716672435),
         (16384, 73695233, 35902597), (32768, 69206017, 3229917)]


def run(plan, n, q, rows):
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randint(0, q, (rows, n), generator=g, device="cuda", dtype=torch.int64).to(plan.dtype)
    b = torch.randint(0, q, (rows, n), generator=g, device="cuda", dtype=torch.int64).to(plan.dtype)
    c = torch.empty_like(a)
    ref = None
    for v, d in plan.variants():
        for _ in range(3):
            tntt.polymul(plan, a, b, out=c, variant=v)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            tntt.polymul(plan, a, b, out=c, variant=v)
        e1.record()
        torch.cuda.synchronize()
        ref = c.clone() if ref is None else ref
        print(d.split()[0], round(rows / (e0.elapsed_time(e1) / 10 * 1e-3) / 1e6, 3), "M/s", d.split("regs=")[1],
              "default" if v == plan.default_variant else "", "same" if torch.equal(ref, c) else "DIFFERENT", flush=True)


for n, q, psi in RINGS:
    plan = tntt.get_plan(n, q, psi, True)
    run(plan, n, q, (128 << 20) // (n * plan.word_bytes))
