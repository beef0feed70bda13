#!/usr/bin/env python3
"""Launch the fused polymul a few times on one configuration (for ncu captures).
usage: run_one.py TAG ROWS [VARIANT] [LAUNCHES]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tiny-ntt_b200"))
import torch  # noqa: E402

import tntt  # noqa: E402
from bench import PARAMS  # noqa: E402

tag = sys.argv[1]
rows = int(sys.argv[2])
variant = int(sys.argv[3]) if len(sys.argv) > 3 and int(sys.argv[3]) >= 0 else None
launches = int(sys.argv[4]) if len(sys.argv) > 4 else 4
p = PARAMS[tag]
plan = tntt.get_plan(p["n"], p["q"], p["psi"], True)
g = torch.Generator(device="cuda").manual_seed(1)
a = torch.randint(0, p["q"], (rows, p["n"]), generator=g, device="cuda", dtype=torch.int64).to(plan.dtype)
b = torch.randint(0, p["q"], (rows, p["n"]), generator=g, device="cuda", dtype=torch.int64).to(plan.dtype)
c = torch.empty_like(a)
for _ in range(launches):
    tntt.polymul(plan, a, b, out=c, variant=variant)
torch.cuda.synchronize()
print("ok", tag, rows, variant, int(c.view(-1)[:8].sum()))
