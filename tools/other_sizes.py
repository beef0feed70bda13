#!/usr/bin/env python3
"""Throughput of the fused / transform-domain kernels at N = 512, 2048, 8192 (JSON lines)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tiny-ntt_b200"))
import torch  # noqa: E402

import tntt  # noqa: E402

Q60 = (1 << 60) - (1 << 14) + 1
RINGS = [(512, 8380417, 1718063), (2048, 8380417, 7901702), (8192, 67043329, 8157893),
         (512, Q60, 984081769261068913), (2048, Q60, 644283108363935541), (8192, Q60, 527760526715669589)]
# rows longer than one CTA: fused product on thread-block clusters only (no transform-domain kernels)
BIG = [(16384, 73695233, 35902597), (32768, 69206017, 3229917),
       (16384, 1152921504606748673, 641000223749548346), (32768, 1152921504606584833, 1100972123716672435)]
for n, q, psi in BIG:
    plan = tntt.get_plan(n, q, psi, True)
    rows = (256 << 20) // (n * plan.word_bytes)
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randint(0, q, (rows, n), generator=g, device="cuda", dtype=torch.int64).to(plan.dtype)
    b = torch.randint(0, q, (rows, n), generator=g, device="cuda", dtype=torch.int64).to(plan.dtype)
    c = torch.empty_like(a)
    for _ in range(3):
        tntt.polymul(plan, a, b, out=c)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        tntt.polymul(plan, a, b, out=c)
    e1.record()
    torch.cuda.synchronize()
    per_s = rows / (e0.elapsed_time(e1) / 10 * 1e-3)
    out = {"n": n, "q_bits": q.bit_length(), "word": plan.word_bytes, "rows": rows, "polymul_per_s": per_s,
           "polymul_GBps": per_s * 3 * n * plan.word_bytes / 1e9,
           "kernel": dict(plan.variants())[plan.default_variant].split()[0]}
    if plan.spectrum:
        spec = tntt.forward_spectrum(plan, b)
        for name, fn in (("forward_spectrum_rows_per_s", lambda: tntt.forward_spectrum(plan, a, out=c)),
                         ("inverse_spectrum_rows_per_s", lambda: tntt.inverse_spectrum(plan, spec, out=c)),
                         ("polymul_spectrum_per_s", lambda: tntt.polymul_spectrum(plan, a, spec, out=c))):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            out[name] = rows / (e0.elapsed_time(e1) / 10 * 1e-3)
        del spec
    print(json.dumps(out), flush=True)
    del a, b, c
    torch.cuda.empty_cache()

for n, q, psi in RINGS:
    plan = tntt.get_plan(n, q, psi, True)
    rows = (512 << 20) // (n * plan.word_bytes)          # 512 MiB per operand: well beyond L2
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randint(0, q, (rows, n), generator=g, device="cuda", dtype=torch.int64).to(plan.dtype)
    b = torch.randint(0, q, (rows, n), generator=g, device="cuda", dtype=torch.int64).to(plan.dtype)
    c = torch.empty_like(a)
    spec = tntt.forward_spectrum(plan, b)

    def t(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return rows / (e0.elapsed_time(e1) / reps * 1e-3)

    out = {"n": n, "q_bits": q.bit_length(), "word": plan.word_bytes, "rows": rows,
           "polymul_per_s": t(lambda: tntt.polymul(plan, a, b, out=c)),
           "polymul_spectrum_per_s": t(lambda: tntt.polymul_spectrum(plan, a, spec, out=c)),
           "forward_spectrum_rows_per_s": t(lambda: tntt.forward_spectrum(plan, a, out=c)),
           "forward_natural_rows_per_s": t(lambda: tntt.forward(plan, a, out=c)),
           "inverse_natural_rows_per_s": t(lambda: tntt.inverse(plan, a, out=c))}
    out["polymul_GBps"] = out["polymul_per_s"] * 3 * n * plan.word_bytes / 1e9
    print(json.dumps(out), flush=True)
    del a, b, c, spec
    torch.cuda.empty_cache()
