#!/usr/bin/env python3
"""Launch time of every fused-polymul variant at small batch sizes (does a shape with more threads per
polynomial win when there are fewer rows than SMs?).  JSON lines on stdout."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tiny-ntt_b200"))
import torch  # noqa: E402

import tntt  # noqa: E402
from bench import PARAMS  # noqa: E402

for tag in sys.argv[1:] or ["n4096_24", "n4096_60"]:
    p = PARAMS[tag]
    plan = tntt.get_plan(p["n"], p["q"], p["psi"], True)
    for rows in (1, 16, 148, 444, 1024):
        g = torch.Generator(device="cuda").manual_seed(rows)
        a = torch.randint(0, p["q"], (rows, p["n"]), generator=g, device="cuda", dtype=torch.int64).to(plan.dtype)
        b = torch.randint(0, p["q"], (rows, p["n"]), generator=g, device="cuda", dtype=torch.int64).to(plan.dtype)
        c = torch.empty_like(a)
        for vid, desc in plan.variants():
            for _ in range(5):
                tntt.polymul(plan, a, b, out=c, variant=vid)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(100):
                tntt.polymul(plan, a, b, out=c, variant=vid)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 10.0
            print(json.dumps({"config": tag, "rows": rows, "variant": vid, "name": desc.split()[0], "us_per_launch": us,
                              "default": vid == plan.default_variant}), flush=True)
