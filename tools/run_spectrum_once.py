#!/usr/bin/env python3
"""Launch the transform-domain kernels a few times on N=4096/60-bit (for ncu captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tiny-ntt_b200"))
import torch, tntt
from bench import PARAMS
tag = sys.argv[1] if len(sys.argv) > 1 else "n4096_60"
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
p = PARAMS[tag]; plan = tntt.get_plan(p["n"], p["q"], p["psi"], True)
g = torch.Generator(device="cuda").manual_seed(1)
a = torch.randint(0, p["q"], (rows, p["n"]), generator=g, device="cuda", dtype=torch.int64).to(plan.dtype)
c = torch.empty_like(a); s = torch.empty_like(a)
for _ in range(3):
    tntt.forward_spectrum(plan, a, out=s); tntt.inverse_spectrum(plan, s, out=c); tntt.polymul_spectrum(plan, a, s, out=c)
torch.cuda.synchronize(); print("ok", bool(torch.equal(tntt.inverse_spectrum(plan, tntt.forward_spectrum(plan, a)), a)))
