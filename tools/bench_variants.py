#!/usr/bin/env python3
"""Time every fused-polymul kernel variant of every shipped parameter set (CUDA events, inputs
larger than L2) and the integer-pipe microbenchmarks.  Writes JSON lines to stdout."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tiny-ntt_b200"))

import torch  # noqa: E402

import tntt  # noqa: E402
from bench import PARAMS, ROWS  # noqa: E402


def main():
    only = sys.argv[1:] or list(PARAMS)
    steps = 10
    for kind, name in enumerate(("imad_lo", "imad_wide_u32", "iadd3", "shoup64_modmul", "shoup32_modmul")):
        print(json.dumps({"microbench": name, "ops_per_s": tntt.microbench(kind)}), flush=True)
    for tag in only:
        p = PARAMS[tag]
        plan = tntt.get_plan(p["n"], p["q"], p["psi"], True)
        rows = ROWS[tag] // 2
        g = torch.Generator(device="cuda").manual_seed(1)
        a = torch.randint(0, p["q"], (rows, p["n"]), generator=g, device="cuda", dtype=torch.int64).to(plan.dtype)
        b = torch.randint(0, p["q"], (rows, p["n"]), generator=g, device="cuda", dtype=torch.int64).to(plan.dtype)
        c = torch.empty_like(a)
        ref = None
        for vid, desc in plan.variants():
            for _ in range(3):
                tntt.polymul(plan, a, b, out=c, variant=vid)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                tntt.polymul(plan, a, b, out=c, variant=vid)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            if ref is None:
                ref = c.clone()
            same = bool(torch.equal(ref, c))
            bytes_ = 3 * p["n"] * plan.word_bytes * rows
            print(json.dumps({"config": tag, "variant": vid, "desc": desc, "rows": rows, "ms": ms,
                              "polymul_per_s": rows / (ms * 1e-3), "GBps": bytes_ / (ms * 1e-3) / 1e9,
                              "agrees_with_first_variant": same}), flush=True)
        # standalone transforms
        spec = tntt.forward_spectrum(plan, b)
        for name, fn in (("forward_spectrum", lambda: tntt.forward_spectrum(plan, a, out=c)),
                         ("inverse_spectrum", lambda: tntt.inverse_spectrum(plan, spec, out=c)),
                         ("polymul_spectrum", lambda: tntt.polymul_spectrum(plan, a, spec, out=c)),
                         ("polymul_spectrum_shared", lambda: tntt.polymul_spectrum(plan, a, spec[0], out=c)),
                         ("forward", lambda: tntt.forward(plan, a, out=c)),
                         ("forward_twist", lambda: tntt.forward(plan, a, twist=True, out=c)),
                         ("inverse_twist", lambda: tntt.inverse(plan, a, twist=True, out=c)),
                         ("pointwise", lambda: tntt.pointwise(plan, a, b, out=c))):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            nb = (3 if name in ("pointwise", "polymul_spectrum") else 2) * p["n"] * plan.word_bytes * rows
            print(json.dumps({"config": tag, "op": name, "rows": rows, "ms": ms, "rows_per_s": rows / (ms * 1e-3),
                              "GBps": nb / (ms * 1e-3) / 1e9}), flush=True)
        del a, b, c, ref
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
