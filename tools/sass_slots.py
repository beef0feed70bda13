#!/usr/bin/env python3
"""Executed-instruction histogram of one kernel from an `ncu --set full --import-source on` report, and the
multiplier-pipe ceiling that follows from it.

    python tools/sass_slots.py gpurun_out/X.ncu-rep [--kernel SUBSTR] [--tag n4096_60] [--variant NAME]
                               [--out profiles/r02_sass_hist_NAME.txt] [--json profiles/sass_slots.json]

The source page of the report carries, per SASS instruction, how many times a warp executed it.  Summed by
opcode that is the DYNAMIC instruction mix (the static `cuobjdump -sass` listing undercounts: ptxas keeps one
copy of code that two inlined calls share).  On sm_100a the integer multiplier lives on the FMA-heavy pipe,
16 lanes wide per SM sub-partition (tools/ubench, profiles/r01_variant_sweep.jsonl microbench lines):

    IMAD.WIDE / IMAD.HI  (32x32 -> 64)   4 pipe-cycles per warp instruction   ("wide",   9.2 T lane-ops/s)
    every other IMAD     (32x32 -> 32)   2 pipe-cycles per warp instruction   ("narrow", 18.2 T lane-ops/s)

so a kernel that executes W wide and M narrow IMADs per warp cannot finish a warp in fewer than 4 W + 2 M
cycles of its sub-partition's multiplier, whatever else it does.  With `warps_per_row` warps per polynomial
pair the ceiling is  SMs * 4 sub-partitions * clock / (warps_per_row * (4 W + 2 M))  polymul/s.
bench.py reads the JSON this writes (profiles/sass_slots.json) for its `roofline` block.
"""
from __future__ import annotations

import argparse
import collections
import csv
import io
import json
import os
import subprocess
import sys


def source_rows(rep: str, kernel: str | None):
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv"]
    if kernel:
        cmd += ["--kernel-name", f"regex:{kernel}"]
    out = subprocess.run(cmd, capture_output=True, text=True, check=True).stdout
    name, header, rows = None, None, []
    for rec in csv.reader(io.StringIO(out)):
        if not rec:
            continue
        if rec[0] == "Kernel Name":
            if name is not None:
                break                      # first matching kernel only
            name = rec[1]
            continue
        if rec[0] == "Address":
            header = rec
            continue
        if header and len(rec) >= len(header) - 2:
            rows.append(dict(zip(header, rec)))
    return name, rows


def classify(op: str) -> str:
    if op.startswith("IMAD.WIDE") or op.startswith("IMAD.HI"):
        return "imad_wide"
    if op.startswith("IMAD.MOV") or op.startswith("IMAD.IADD") or op.startswith("IMAD.SHL") or op.startswith("IMAD.X"):
        return "imad_narrow_nonproduct"
    if op.startswith("IMAD"):
        return "imad_narrow"
    if op.startswith(("LDG", "STG", "LD.", "ST.", "LDL", "STL", "CCTL")):
        return "global_local_mem"
    if op.startswith(("LDS", "STS")):
        return "shared_mem"
    if op.startswith(("BAR", "BRA", "EXIT", "WARPSYNC", "BSSY", "BSYNC", "NOP", "CALL", "RET")):
        return "control"
    if op.startswith(("LDC", "ULDC", "S2R", "S2UR", "CS2R", "MOV", "UMOV", "R2UR")) or op.startswith("U"):
        return "uniform_const_move"
    return "alu"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--kernel", default=None, help="regex on the kernel name (default: first kernel in the report)")
    ap.add_argument("--warps-per-row", type=float, default=8, help="warps that share one polynomial pair (N=256: 16 lanes = 0.5)")
    ap.add_argument("--sms", type=int, default=148)
    ap.add_argument("--clock-mhz", type=float, default=1965.0)
    ap.add_argument("--tag", default=None)
    ap.add_argument("--variant", default=None)
    ap.add_argument("--out", default=None)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()

    name, rows = source_rows(a.report, a.kernel)
    if not rows:
        sys.exit("no source rows: was the report captured with --import-source on / -lineinfo?")
    per_op, warps = collections.Counter(), 0
    for r in rows:
        src = r["Source"].strip()
        toks = src.split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        op = op.rstrip(";")
        n = int(float(r.get("Instructions Executed", "0") or 0))
        per_op[op] += n
        if warps == 0:
            warps = n                      # the entry instruction runs once per warp
    total = sum(per_op.values())
    cat = collections.Counter()
    for op, n in per_op.items():
        cat[classify(op)] += n
    per_warp = {k: v / warps for k, v in cat.items()}
    wide = per_warp.get("imad_wide", 0.0)
    narrow = per_warp.get("imad_narrow", 0.0) + per_warp.get("imad_narrow_nonproduct", 0.0)
    pipe_cycles = 4 * wide + 2 * narrow
    ceiling = a.sms * 4 * a.clock_mhz * 1e6 / (a.warps_per_row * pipe_cycles)
    issue_ceiling = a.sms * 4 * a.clock_mhz * 1e6 / (a.warps_per_row * total / warps)
    lines = [
        f"kernel: {name}",
        f"report: {os.path.basename(a.report)}   warps launched: {warps}   warp instructions executed: {total}",
        f"per warp: {total / warps:.1f} instructions",
        "",
        "category                    per warp   share",
    ]
    for k, v in sorted(per_warp.items(), key=lambda kv: -kv[1]):
        lines.append(f"  {k:<24} {v:9.1f}   {v * warps / total:6.1%}")
    lines += [
        "",
        f"multiplier pipe: {wide:.1f} wide x 4 + {narrow:.1f} narrow x 2 = {pipe_cycles:.0f} cycles per warp",
        f"ceiling (multiplier pipe): {a.sms} SMs x 4 x {a.clock_mhz:.0f} MHz / ({a.warps_per_row} warps x {pipe_cycles:.0f}) = {ceiling / 1e6:.2f} M polymul/s",
        f"ceiling (one instruction issued per cycle per sub-partition): {issue_ceiling / 1e6:.2f} M polymul/s",
        "",
        "opcode                              executed   per warp",
    ]
    for op, n in per_op.most_common():
        if n:
            lines.append(f"  {op:<32} {n:>10}   {n / warps:8.2f}")
    text = "\n".join(lines) + "\n"
    print(text)
    if a.out:
        with open(a.out, "w") as fh:
            fh.write(text)
    if a.json and a.tag:
        db = {}
        if os.path.exists(a.json):
            with open(a.json) as fh:
                db = json.load(fh)
        db[a.tag] = {
            "variant": a.variant or name, "warps_per_row": a.warps_per_row, "inst_per_warp": total / warps,
            "imad_wide_per_warp": wide, "imad_narrow_per_warp": narrow, "pipe_cycles_per_warp": pipe_cycles,
            "source": os.path.basename(a.out) if a.out else os.path.basename(a.report),
        }
        with open(a.json, "w") as fh:
            json.dump(db, fh, indent=1, sort_keys=True)
            fh.write("\n")


if __name__ == "__main__":
    main()
