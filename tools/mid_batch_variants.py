#!/usr/bin/env python3
"""Every fused variant of one parameter set at mid-size batches (BASELINE config 2 is 2^16 rows of N = 256: nine waves of
CTAs, where launch ramp and tail are a tenth of the launch).  usage: mid_batch_variants.py [TAG] [ROWS ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tiny-ntt_b200"))

import torch  # noqa: E402

import tntt  # noqa: E402
from bench import PARAMS  # noqa: E402


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "dilithium"
    sizes = [int(v) for v in sys.argv[2:]] or [1 << 16, 1 << 17, 1 << 18]
    p = PARAMS[tag]
    plan = tntt.get_plan(p["n"], p["q"], p["psi"], True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for rows in sizes:
        g = torch.Generator(device="cuda").manual_seed(1)
        a = torch.randint(0, p["q"], (rows, p["n"]), generator=g, device="cuda", dtype=torch.int64).to(plan.dtype)
        b = torch.randint(0, p["q"], (rows, p["n"]), generator=g, device="cuda", dtype=torch.int64).to(plan.dtype)
        c = torch.empty_like(a)
        for vid, desc in plan.variants():
            if "_c" in desc.split(" ")[0].split("_b")[-1]:
                continue                     # cluster shapes: small batches only
            for _ in range(3):
                tntt.polymul(plan, a, b, out=c, variant=vid)
            best, tot, reps = 1e9, 0.0, 20
            for _ in range(reps):            # one launch per measurement, L2 flushed in between
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                tntt.polymul(plan, a, b, out=c, variant=vid)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                best, tot = min(best, ms), tot + ms
            print(json.dumps({"config": tag, "rows": rows, "variant": vid, "desc": desc.split(" ")[0], "default": vid == plan.default_variant,
                              "us_mean": 1e3 * tot / reps, "us_best": 1e3 * best, "polymul_per_s_mean": rows / (tot / reps * 1e-3)}), flush=True)


if __name__ == "__main__":
    main()
