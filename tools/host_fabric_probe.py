#!/usr/bin/env python3
"""What can the HOST side of this box sink?  One process, G GPUs, pinned buffers, one copy stream pair per GPU.

    python tools/host_fabric_probe.py [--gpus G] [--mib 256] [--reps 8]

For g = 1, 2, 4, ... G GPUs copying at the same time it reports the aggregate and per-GPU GB/s of
  h2d   : host -> device only
  d2h   : device -> host only
  mix21 : the 2:1 mix of the polymul path (two operand rows in, one product row out), both directions concurrently
and converts the mix into the end-to-end ceiling of the N = 4096 / 60-bit polymul (98 304 bytes over PCIe per product).
The numbers bound tntt_polymul_host / tntt_polymul_host_multi from above whatever the GPUs do (task 4).
"""
import argparse
import json
import time

import torch

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=0)
ap.add_argument("--mib", type=int, default=256)
ap.add_argument("--reps", type=int, default=8)
args = ap.parse_args()
G = args.gpus or torch.cuda.device_count()
n = args.mib << 20

dev_in = [torch.empty(2 * n, dtype=torch.uint8, device=f"cuda:{g}") for g in range(G)]
dev_out = [torch.empty(n, dtype=torch.uint8, device=f"cuda:{g}") for g in range(G)]
host_in = [torch.empty(2 * n, dtype=torch.uint8).pin_memory() for _ in range(G)]
host_out = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(G)]
for h in host_in + host_out:
    h.fill_(1)
s_in = [torch.cuda.Stream(device=g) for g in range(G)]
s_out = [torch.cuda.Stream(device=g) for g in range(G)]


def sync(gs):
    for g in gs:
        torch.cuda.synchronize(g)


def run(gs, h2d, d2h):
    def once():
        for g in gs:
            if h2d:
                with torch.cuda.stream(s_in[g]):
                    dev_in[g].copy_(host_in[g], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s_out[g]):
                    host_out[g].copy_(dev_out[g], non_blocking=True)
    once()
    sync(gs)
    t = time.perf_counter()
    for _ in range(args.reps):
        once()
    sync(gs)
    dt = time.perf_counter() - t
    moved_in = 2 * n * args.reps * len(gs) if h2d else 0
    moved_out = n * args.reps * len(gs) if d2h else 0
    return moved_in / dt / 1e9, moved_out / dt / 1e9


g = 1
counts = []
while g < G:
    counts.append(g)
    g *= 2
counts.append(G)
for g in counts:
    gs = list(range(g))
    rec = {"gpus": g, "mib_per_copy": args.mib}
    a, _ = run(gs, True, False)
    rec["h2d_total_gbs"], rec["h2d_per_gpu_gbs"] = a, a / g
    _, b = run(gs, False, True)
    rec["d2h_total_gbs"], rec["d2h_per_gpu_gbs"] = b, b / g
    a, b = run(gs, True, True)
    rec["mix21_h2d_total_gbs"], rec["mix21_d2h_total_gbs"] = a, b
    rec["mix21_total_gbs"] = a + b
    rec["e2e_ceiling_polymul_per_s_n4096_60"] = (a + b) * 1e9 / 98304
    print(json.dumps(rec), flush=True)
