#!/usr/bin/env python3
"""Launch the default, the one-CTA-per-SM and the cluster shape of the N=4096 polymul at small batch sizes,
a few launches each, so that `ncu --metrics gpu__time_duration.sum` reports exact kernel durations (CUDA-event
timing of back-to-back launches is quantised to the ~2 us completion-to-launch cadence).
usage: small_batch_launches.py [TAG]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tiny-ntt_b200"))
import torch  # noqa: E402

import tntt  # noqa: E402
from bench import PARAMS  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "n4096_60"
p = PARAMS[tag]
plan = tntt.get_plan(p["n"], p["q"], p["psi"], True)
chosen = [v for v in (plan.default_variant, plan.small_variant, plan.cluster_variant) if v >= 0]
chosen += [v for v, d in plan.variants() if d.split()[0].endswith("_c4") and v not in chosen]
for rows in (1, 16, 37, 148):
    g = torch.Generator(device="cuda").manual_seed(rows)
    a = torch.randint(0, p["q"], (rows, p["n"]), generator=g, device="cuda", dtype=torch.int64).to(plan.dtype)
    b = torch.randint(0, p["q"], (rows, p["n"]), generator=g, device="cuda", dtype=torch.int64).to(plan.dtype)
    c = torch.empty_like(a)
    for v in chosen:
        for _ in range(3):
            tntt.polymul(plan, a, b, out=c, variant=v)
        torch.cuda.synchronize()
print("ok")
