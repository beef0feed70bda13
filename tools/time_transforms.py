#!/usr/bin/env python3
"""Time the standalone transforms / polymul with a chosen libtntt build (TNTT_LIB env)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tiny-ntt_b200"))
import torch
import tntt._lib as L
if os.environ.get("TNTT_LIB"):
    L.LIB_PATH = os.environ["TNTT_LIB"]
import tntt
from bench import PARAMS, ROWS
for tag in sys.argv[1:] or list(PARAMS):
    p = PARAMS[tag]; plan = tntt.get_plan(p["n"], p["q"], p["psi"], True)
    rows = ROWS[tag] // 2
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randint(0, p["q"], (rows, p["n"]), generator=g, device="cuda", dtype=torch.int64).to(plan.dtype)
    b = a.clone(); c = torch.empty_like(a)
    for name, fn in (("forward", lambda: tntt.forward(plan, a, out=c)), ("inverse", lambda: tntt.inverse(plan, a, out=c)),
                     ("polymul", lambda: tntt.polymul(plan, a, b, out=c))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): fn()
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 10)
        print(os.environ.get("TNTT_LIB", "default")[-8:], tag, name, "%.3f ms" % best, "%.1fM rows/s" % (rows / best / 1e3), flush=True)
