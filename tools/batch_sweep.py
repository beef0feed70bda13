#!/usr/bin/env python3
"""BASELINE config 5: N=4096, 24-bit modulus, batch-size sweep vs the reference's AVX-512 build on the host.
Prints one JSON line per batch size (GPU polymul/s by CUDA events, device-resident data) and one for the CPU."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tiny-ntt_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import tntt  # noqa: E402
from bench import PARAMS, cpu_reference_throughput  # noqa: E402


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "n4096_24"
    p = PARAMS[tag]
    plan = tntt.get_plan(p["n"], p["q"], p["psi"], True)
    for logb in (0, 4, 8, 12, 16):
        rows = 1 << logb
        g = torch.Generator(device="cuda").manual_seed(logb)
        a = torch.randint(0, p["q"], (rows, p["n"]), generator=g, device="cuda", dtype=torch.int64).to(plan.dtype)
        b = torch.randint(0, p["q"], (rows, p["n"]), generator=g, device="cuda", dtype=torch.int64).to(plan.dtype)
        c = torch.empty_like(a)
        for _ in range(5):
            tntt.polymul(plan, a, b, out=c)
        torch.cuda.synchronize()
        reps = 200 if rows <= 4096 else 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            tntt.polymul(plan, a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(json.dumps({"config": tag, "batch": rows, "us_per_launch": ms * 1e3, "polymul_per_s": rows / (ms * 1e-3),
                          "note": "inputs fit L2 below batch 4096" if rows < 4096 else "inputs exceed L2"}), flush=True)
    ctx = cpu_reference_throughput(tag, seconds=5.0)
    t = time.perf_counter()
    ctx["run"](ctx["a"], ctx["b"])
    dt = time.perf_counter() - t
    one = ctx["a"][:1], ctx["b"][:1]
    from oracle.cpu_ref import RefLib
    lat = None
    if RefLib.available(tag):
        lib = RefLib(tag)
        t = time.perf_counter()
        for _ in range(50):
            lib.polymul(one[0], one[1], threads=1)
        lat = (time.perf_counter() - t) / 50 * 1e6
    print(json.dumps({"config": tag, "cpu": ctx["name"], "host_threads": ctx["cores"], "polymul_per_s": ctx["rows"] / dt,
                      "single_polymul_latency_us_1_thread": lat}), flush=True)


if __name__ == "__main__":
    main()
