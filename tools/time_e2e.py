#!/usr/bin/env python3
"""End-to-end (pinned host buffers) polymul throughput for several pipeline chunk sizes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tiny-ntt_b200"))
import torch
import tntt
from bench import PARAMS
p = PARAMS["n4096_60"]; plan = tntt.get_plan(p["n"], p["q"], p["psi"], True)
rows = 8192
a = torch.randint(0, p["q"], (rows, p["n"]), dtype=torch.int64).pin_memory()
b = torch.randint(0, p["q"], (rows, p["n"]), dtype=torch.int64).pin_memory()
c = torch.empty_like(a).pin_memory()
for mb in (2, 4, 8, 16, 32, 64, 128):
    os.environ["TNTT_HOST_CHUNK_MB"] = str(mb)
    tntt.polymul_host(plan, a, b, out=c)
    t = time.perf_counter()
    for _ in range(5): tntt.polymul_host(plan, a, b, out=c)
    dt = (time.perf_counter() - t) / 5
    print(f"chunk {mb} MiB: {dt*1e3:.2f} ms/step  {rows/dt/1e3:.0f}k polymul/s  H2D {2*rows*32768/dt/1e9:.1f} GB/s")
