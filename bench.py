#!/usr/bin/env python3
"""Headline benchmark: batched negacyclic polymul, N = 4096, 60-bit modulus (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--rows R] [--config TAG]

Own arm: one process per GPU (torchrun for N > 1).  The batch is sharded by rows, there is no
data-path collective; torch.distributed only provides the barrier and the
max-over-ranks reduction of the device time.  A "step" is one fused polymul launch over the
rank's whole shard, inputs resident in HBM.  One JSON line is printed by rank 0.  Beside the headline
it carries (outside the headline's timed region): a >= 2 s sustained repeat with its own clock samples,
BASELINE.json's other configurations (`other_configs`: Dilithium 2^16 / 2^20 rows, N=1024 24-bit 2^18 rows,
the N=4096 24-bit batch-size sweep), the end-to-end leg through host buffers, and the CPU baseline.

Reference arm (--impl reference): the reference's own C++ implementation of the path
(software_benchmark/benchmark_ntt_60bit.cpp compiled unmodified into oracle/_ref) on all host
threads, rank 0 only, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tiny-ntt_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "polymuls_per_sec_n4096_60bit"
UNIT = "polymul/s"

# workloads: tag -> rows per GPU (inputs exceed the 126 MB L2 in every case)
ROWS = {"n4096_60": 1 << 15, "n4096_24": 1 << 16, "n1024_24": 1 << 18, "dilithium": 1 << 20}
PARAMS = {
    "dilithium": dict(n=256, q=8380417, psi=1239911),
    "n1024_24": dict(n=1024, q=8380417, psi=5548360),
    "n4096_24": dict(n=4096, q=8380417, psi=283817),
    "n4096_60": dict(n=4096, q=(1 << 60) - (1 << 14) + 1, psi=431606828070683274),
}
# algorithmic work per polymul
MODMULS = {256: 3584, 1024: 17408, 4096: 81920}
IMAD_PER_MODMUL = {4: 3.0, 8: 10.05}   # u64: (73728*10 + 4096*11 + 4096*10) / 81920
# BASELINE.json configs 2, 3 and 5, timed after the headline (rows per GPU; the sweep's batch is the TOTAL batch)
OTHER_FIXED = (("dilithium", 1 << 16), ("dilithium", 1 << 20), ("n1024_24", 1 << 18))
SWEEP_TAG, SWEEP_BATCHES = "n4096_24", (1, 1 << 4, 1 << 8, 1 << 12, 1 << 16)
PARITY_ROWS = 256                      # >= 256 random rows element-wise before timing


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--config", default="n4096_60", choices=sorted(PARAMS))
    ap.add_argument("--rows", type=int, default=0, help="rows per GPU (default: workload table)")
    ap.add_argument("--variant", type=int, default=-1, help="force a kernel variant (benchmarking)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--sustained-seconds", type=float, default=2.0)
    return ap.parse_args()


def workload_config(tag, rows_per_gpu, gpus):
    """The `config` object: what is computed, identical in both arms (arm-specific detail goes under `run`)."""
    p = PARAMS[tag]
    wb = 8 if p["q"] >> 27 else 4
    return {"workload": workload_name(tag), "n": p["n"], "q": p["q"], "psi": p["psi"], "rows_per_gpu": rows_per_gpu,
            "rows_total": rows_per_gpu * gpus, "gpus": gpus,
            "l2": "inputs exceed L2 (%.2f GB read per launch per GPU vs 126 MB)" % (2 * p["n"] * wb * rows_per_gpu / 1e9)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING the timed region.

    NVML is queried in-process from a thread (one light call per quantity every 20 ms).  Spawning `nvidia-smi -lms`
    for this stalls the GPU for a few milliseconds per query, which is visible in a 50 ms timed region; it is
    only the fallback when the NVML binding is missing."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.nvml, self.stop_flag = index, [], None, None, False
        self.thread = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def start(self):
        try:
            if os.environ.get("TNTT_BENCH_SAMPLER") == "smi":
                raise RuntimeError("nvidia-smi sampler forced")
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                try:
                    watts = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                except Exception:
                    watts = float("nan")
                self.rows.append((time.time(), sm, watts, mask))
            except Exception:
                pass
            time.sleep(0.02)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        names = [r[0] for r in self.REASONS]
        if self.nvml:
            time.sleep(0.03)
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            rows = [r for r in self.rows if t0 - 0.02 <= r[0] <= t1 + 0.03] or self.rows
            sm = sorted(r[1] for r in rows)
            power = [r[2] for r in rows if r[2] == r[2]]
            reasons = sorted({name for r in rows for name, bit in self.REASONS if r[3] & bit})
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_sm, "reasons": reasons,
                    "samples": len(sm), "power_w_max": max(power) if power else None, "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, power = [], None, set(), []
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for (_, r) in self.rows]
        for line in rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
                power.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None, "source": "nvidia-smi"}


def cpu_reference_throughput(tag: str, seconds: float = 10.0):
    """polymul/s of the reference's own C++ code on all host threads, on a bounded sample."""
    import numpy as np

    from oracle.cpu_ref import COracle, RefLib, best_simd

    p = PARAMS[tag]
    cores = os.cpu_count() or 1
    try:
        affinity = len(os.sched_getaffinity(0))
        cores = min(cores, affinity)
    except AttributeError:
        pass
    rng = np.random.default_rng(1234)
    if RefLib.available(tag):
        lib = RefLib(tag)
        kind, name = "reference", f"oracle/_ref/libref_{tag}_{lib.simd}.so (reference C++ sources, {lib.simd})"
        run = lambda a, b: lib.polymul(a, b, threads=cores)          # noqa: E731
        dt = lib.dtype
    else:
        co = COracle()
        kind, name = "port", "oracle/ntt_oracle.c (C restatement)"
        run = lambda a, b: co.nwc_poly_mult(a, b, p["psi"], p["q"], threads=cores)  # noqa: E731
        dt = np.uint64
    probe = max(cores * 2, 8)
    a = rng.integers(0, p["q"], size=(probe, p["n"]), dtype=np.uint64).astype(dt)
    b = rng.integers(0, p["q"], size=(probe, p["n"]), dtype=np.uint64).astype(dt)
    run(a, b)
    t = time.perf_counter()
    run(a, b)
    per_row = (time.perf_counter() - t) / probe
    rows = int(max(probe, min(seconds / max(per_row, 1e-9), 1 << 17)))
    a = rng.integers(0, p["q"], size=(rows, p["n"]), dtype=np.uint64).astype(dt)
    b = rng.integers(0, p["q"], size=(rows, p["n"]), dtype=np.uint64).astype(dt)
    return dict(run=run, a=a, b=b, rows=rows, cores=cores, kind=kind, name=name)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    tag = args.config
    ctx = cpu_reference_throughput(tag, seconds=max(1.0, min(8.0, 120.0 / max(1, args.steps + args.warmup))))
    for _ in range(args.warmup):
        ctx["run"](ctx["a"], ctx["b"])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx["run"](ctx["a"], ctx["b"])
    dt = time.perf_counter() - t0
    value = ctx["rows"] * args.steps / dt
    p = PARAMS[tag]
    sample = f"{ctx['rows']} polymuls per step on {ctx['cores']} host threads, {ctx['name']}"
    line = {
        "impl": "reference", "metric": METRIC if tag == "n4096_60" else f"polymuls_per_sec_{tag}", "value": value,
        "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64" if p["q"] >> 32 else "u32", "data": "synthetic",
        "config": workload_config(tag, args.rows or ROWS[tag], args.gpus),
        "run": {"host_threads": ctx["cores"], "implementation": ctx["name"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": ctx["cores"], "kind": ctx["kind"], "sample": sample,
                         "sample_rows": ctx["rows"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(tag):
    p = PARAMS[tag]
    return (f"batched negacyclic polymul N={p['n']} q={p['q']} ({p['q'].bit_length()}-bit) psi={p['psi']} "
            f"(forward NTT x2 -> pointwise -> inverse NTT), rows sharded across GPUs")


def load_json(name):
    path = os.path.join(ROOT, "profiles", name)
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh)
    return {}


def parity_gate(tntt, plan, a, b, c, tag, rows):
    """the first `rows` (random) rows of the batch, element-wise against the reference's
    own C++ code (oracle/_ref) or, where that was not built, the C restatement.  Row 0/1 are the C++ benchmark's LCG
    polynomials, whose product checksum is a golden constant of the reference."""
    import numpy as np
    import torch

    from oracle.cpu_ref import COracle, RefLib
    from tntt import fixtures

    p = PARAMS[tag]
    n, q, psi = p["n"], p["q"], p["psi"]
    wb = plan.word_bytes
    npdt, sdt = (np.uint32, np.int32) if wb == 4 else (np.uint64, np.int64)
    a[0] = torch.from_numpy(np.array(fixtures.make_poly(1, n, q), dtype=npdt).view(sdt)).cuda()
    b[0] = torch.from_numpy(np.array(fixtures.make_poly(2, n, q), dtype=npdt).view(sdt)).cuda()
    tntt.polymul(plan, a, b, out=c)
    torch.cuda.synchronize()
    rows = min(rows, a.shape[0])
    ha, hb = a[:rows].cpu().numpy().view(npdt), b[:rows].cpu().numpy().view(npdt)
    got = c[:rows].cpu().numpy().view(npdt)
    if RefLib.available(tag):
        lib = RefLib(tag)
        want, checker = lib.polymul(ha.astype(lib.dtype), hb.astype(lib.dtype), threads=os.cpu_count() or 1), f"oracle/_ref ({lib.simd})"
    else:
        want, checker = COracle().nwc_poly_mult(ha.astype(np.uint64), hb.astype(np.uint64), psi, q, threads=os.cpu_count() or 1), "oracle/ntt_oracle.c"
    if not (got.astype(np.uint64) == np.asarray(want).astype(np.uint64)).all():
        raise SystemExit(f"parity gate failed: {rows} rows differ from {checker}")
    golden = fixtures.REFERENCE_CHECKSUMS[(n, q)]
    if fixtures.checksum(got[0].tolist(), q) != golden:
        raise SystemExit(f"parity gate failed: checksum of row 0 != reference {golden}")
    return {"rows": rows, "checker": checker, "row0_checksum": golden}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import tntt
    from tntt.shard import max_over_ranks

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    from tntt.shard import bind_host_thread_to_gpu
    all_cpus = os.sched_getaffinity(0)
    numa = bind_host_thread_to_gpu(local)   # staging buffers of the e2e leg stay next to this rank's GPU
    if world > 1:
        # NCCL prints its version banner on stdout when the first communicator is created; rank 0 must print
        # exactly one JSON line, so stdout is parked on /dev/null while the communicator comes up.
        sys.stdout.flush()
        saved, devnull = os.dup(1), os.open(os.devnull, os.O_WRONLY)
        os.dup2(devnull, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved, 1)
            os.close(saved)
            os.close(devnull)
    tag = args.config
    p = PARAMS[tag]
    n, q, psi = p["n"], p["q"], p["psi"]
    if args.variant >= 0:      # a private, uncached plan: set_default_variant rewrites the plan's dispatch fields
        plan = tntt.Plan.create(n, q, psi, True, local)
        plan.set_default_variant(args.variant)
    else:
        plan = tntt.get_plan(n, q, psi, True, local)
    variant_desc = dict(plan.variants()).get(plan.default_variant, "literal-schedule path")
    variant_name = variant_desc.split(" ")[0]
    rows = args.rows or ROWS[tag]
    wb = plan.word_bytes
    warmup = max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def operands(pl, nrows, seed):
        gen = torch.Generator(device="cuda").manual_seed(seed)
        x = torch.randint(0, pl.q, (nrows, pl.n), generator=gen, device="cuda", dtype=torch.int64).to(pl.dtype)
        y = torch.randint(0, pl.q, (nrows, pl.n), generator=gen, device="cuda", dtype=torch.int64).to(pl.dtype)
        return x, y, torch.empty_like(x)

    def timed(fn, reps, warm=3):
        """reps launches between two events on the launching stream, barrier + synchronize on both sides, max over ranks"""
        for _ in range(warm):
            fn()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(reps):
            fn()
        s1.record()
        barrier()
        return max_over_ranks(s0.elapsed_time(s1))

    a, b, c = operands(plan, rows, 1234 + rank)
    gate = parity_gate(tntt, plan, a, b, c, tag, PARITY_ROWS)

    # ---- the headline: K launches over the rank's shard, inputs resident in HBM ------------------------------
    for _ in range(warmup):
        tntt.polymul(plan, a, b, out=c)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for _ in range(args.steps):
        tntt.polymul(plan, a, b, out=c)
    e1.record()
    barrier()
    t1 = time.time()
    ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    total_rows = rows * world
    value = total_rows * args.steps / (ms * 1e-3)
    launches = args.steps

    # ---- the same launch repeated for >= 2 s: sustained clocks and power ("57 ms timed region") ----
    sustained = None
    if args.sustained_seconds > 0:
        per_launch = ms * 1e-3 / args.steps
        reps = max(args.steps, int(args.sustained_seconds / per_launch) + 1)
        sampler2 = ClockSampler(local)
        if rank == 0:
            sampler2.start()
            time.sleep(0.1)
        ts0 = time.time()
        s_ms = timed(lambda: tntt.polymul(plan, a, b, out=c), reps, warm=0)
        ts1 = time.time()
        sustained = {"value": total_rows * reps / (s_ms * 1e-3), "unit": UNIT, "launches": reps, "seconds": s_ms * 1e-3,
                     "clocks": sampler2.stop(ts0, ts1) if rank == 0 else None}
        launches += reps

    # ---- the same work with operands kept in the transform domain (reported next to the headline)
    extras = None
    if plan.spectrum:
        reps = max(3, min(args.steps, 10))
        rate = lambda fn: total_rows * reps / (timed(fn, reps, warm=2) * 1e-3)   # noqa: E731
        spec = tntt.forward_spectrum(plan, b)
        extras = {
            "forward_spectrum_rows_per_s": rate(lambda: tntt.forward_spectrum(plan, a, out=c)),
            "inverse_spectrum_rows_per_s": rate(lambda: tntt.inverse_spectrum(plan, spec, out=c)),
            "polymul_spectrum_per_s": rate(lambda: tntt.polymul_spectrum(plan, a, spec, out=c)),
            "pointwise_rows_per_s": rate(lambda: tntt.pointwise(plan, a, spec, out=c)),
            "note": "one operand (or both) kept as a spectrum: tntt_spectrum_forward / tntt_polymul_spectrum / "
                    "tntt_pointwise + tntt_spectrum_inverse; device-resident, whole job over all GPUs",
        }
        launches += 4 * (reps + 2) + 1
        tntt.polymul(plan, a, b, out=c)      # restore c for the e2e parity check below
        del spec

    # ---- end to end through the host-buffer entry points (pinned memory, H2D + kernel + D2H per step) ------
    e2e = None
    if not args.no_e2e:
        # one step = the rank's whole shard, as in the device-resident measurement (N = 1: 3 x 1.07 GB of pinned buffers);
        # under torchrun a quarter of it per GPU, so that rank 0's single-process arena for all GPUs stays small
        e_rows = rows if world == 1 else min(rows, max(1, (256 << 20) // (n * wb)))
        bytes_row = n * wb
        e_steps = max(3, min(args.steps, 10))
        ha = a[:e_rows].cpu().pin_memory()
        hb = b[:e_rows].cpu().pin_memory()
        hc = torch.empty_like(ha).pin_memory()
        tntt.polymul_host(plan, ha, hb, out=hc)
        if not torch.equal(hc, c[:e_rows].cpu()):
            raise SystemExit("e2e parity failed: host pipeline result differs from the device result")
        barrier()
        tt0 = time.perf_counter()
        for _ in range(e_steps):
            tntt.polymul_host(plan, ha, hb, out=hc)       # blocking: returns when hc is complete
        torch.cuda.synchronize()
        e_ms = max_over_ranks((time.perf_counter() - tt0) * 1e3)
        e2e = {"value": e_rows * world * e_steps / (e_ms * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": 2 * e_rows * bytes_row * world, "d2h_bytes_per_step": e_rows * bytes_row * world,
               "rows_per_step_per_gpu": e_rows, "steps": e_steps, "ms_per_step": e_ms / e_steps,
               "pcie_gbs_per_gpu": 3 * e_rows * bytes_row * e_steps / (e_ms * 1e-3) / 1e9,
               "api": "tntt_polymul_host, one process per GPU (pinned host buffers; H2D / kernel / D2H pipelined over 4 "
                      "streams, ramped chunks)"}
        launches += (e_steps + 1) * 8
        # the same step with the second operand cached on the device as a spectrum (the fixed-key pattern the reference's
        # report targets, reports/final-report.tex:571-610): 1/3 less PCIe traffic, both directions equally loaded.  A
        # side figure: the headline e2e above moves BOTH operands every step.
        if plan.info.spectrum:
            spec = tntt.forward_spectrum(plan, b[:e_rows])
            tntt.polymul_spectrum_host(plan, ha, spec, out=hc)
            if not torch.equal(hc, c[:e_rows].cpu()):
                raise SystemExit("e2e parity failed: cached-operand host pipeline result differs from the device result")
            barrier()
            tt0 = time.perf_counter()
            for _ in range(e_steps):
                tntt.polymul_spectrum_host(plan, ha, spec, out=hc)
            torch.cuda.synchronize()
            k_ms = max_over_ranks((time.perf_counter() - tt0) * 1e3)
            e2e["cached_operand"] = {
                "value": e_rows * world * e_steps / (k_ms * 1e-3), "unit": UNIT, "ms_per_step": k_ms / e_steps,
                "h2d_bytes_per_step": e_rows * bytes_row * world, "d2h_bytes_per_step": e_rows * bytes_row * world,
                "pcie_gbs_per_gpu": 2 * e_rows * bytes_row * e_steps / (k_ms * 1e-3) / 1e9,
                "api": "tntt_polymul_spectrum_host: a from pinned host memory, b kept on the device as spectra "
                       "(tntt_spectrum_forward, one per row), c to pinned host memory"}
            launches += (e_steps + 1) * 8
            del spec
        if world > 1:
            # the single-process form: rank 0 drives every GPU through ONE call on one pinned
            # arena (tntt_polymul_host_multi: per-device plan + streams, host barrier); the other ranks stay idle
            del ha, hb, hc
            barrier()
            single = None
            if rank == 0:
                os.sched_setaffinity(0, all_cpus)
                plans = tntt.get_plans(n, q, psi, True, list(range(world)))
                ga = a[:e_rows].cpu().repeat(world, 1).pin_memory()
                gb = b[:e_rows].cpu().repeat(world, 1).pin_memory()
                gc = torch.empty_like(ga).pin_memory()
                tntt.polymul_sharded(plans, ga, gb, out=gc)
                if not torch.equal(gc[-e_rows:], c[:e_rows].cpu()):
                    raise SystemExit("e2e parity failed: sharded host pipeline result differs from the device result")
                tt0 = time.perf_counter()
                for _ in range(e_steps):
                    tntt.polymul_sharded(plans, ga, gb, out=gc)
                s_ms = (time.perf_counter() - tt0) * 1e3
                single = {"value": e_rows * world * e_steps / (s_ms * 1e-3), "unit": UNIT, "ms_per_step": s_ms / e_steps,
                          "pcie_gbs_per_gpu": 3 * e_rows * bytes_row * e_steps / (s_ms * 1e-3) / 1e9,
                          "api": "tntt_polymul_host_multi, ONE process driving all GPUs (one pinned arena, a host thread "
                                 "per device, host barrier)"}
                del ga, gb, gc
            barrier()
            if rank == 0:
                e2e["one_process_per_gpu"] = {k: e2e[k] for k in ("value", "ms_per_step", "pcie_gbs_per_gpu", "api")}
                e2e["single_process"] = single
                if single["value"] > e2e["value"]:      # the headline e2e is the better of the two public entry points
                    e2e.update({k: single[k] for k in ("value", "ms_per_step", "pcie_gbs_per_gpu", "api")})

    # ---- BASELINE.json configs 2, 3, 5 (outside every timed region above) ------------------------------------
    others = None
    if not args.no_other_configs and tag == "n4096_60":
        del a, b, c
        torch.cuda.empty_cache()
        others = []
        from tntt import fixtures
        traffic_db = load_json("traffic.json")
        peaks, _ = measured_peaks()
        peak_lo = None
        try:
            peak_lo = tntt.microbench(0, local)
        except Exception:
            pass

        def one(otag, nrows_gpu, batch_total=None):
            nonlocal launches
            op = PARAMS[otag]
            pl = tntt.get_plan(op["n"], op["q"], op["psi"], True, local)
            x, y, z = operands(pl, max(nrows_gpu, 1), 99 + rank)
            active = nrows_gpu > 0
            if active and rank == 0:      # row 0 = the C++ benchmark's LCG polynomials: the reference's `checksum=` constant
                odt, osd = (np.uint32, np.int32) if pl.word_bytes == 4 else (np.uint64, np.int64)
                x[0] = torch.from_numpy(np.array(fixtures.make_poly(1, op["n"], op["q"]), dtype=odt).view(osd)).cuda()
                y[0] = torch.from_numpy(np.array(fixtures.make_poly(2, op["n"], op["q"]), dtype=odt).view(osd)).cuda()
                tntt.polymul(pl, x[:1], y[:1], out=z[:1])
                if fixtures.checksum(z[0].cpu().numpy().view(odt).tolist(), op["q"]) != fixtures.REFERENCE_CHECKSUMS[(op["n"], op["q"])]:
                    raise SystemExit(f"parity gate failed for {otag}: row-0 checksum differs from the reference binary's")
            fn = (lambda: tntt.polymul(pl, x[:nrows_gpu], y[:nrows_gpu], out=z[:nrows_gpu])) if active else (lambda: None)
            per = max(nrows_gpu, 1) * op["n"] * pl.word_bytes * 3
            reps = int(max(5, min(400, 2e9 / per)))
            o_ms = timed(fn, reps, warm=3)
            launches += reps + 3
            tot = batch_total if batch_total is not None else nrows_gpu * world
            val = tot * reps / (o_ms * 1e-3)
            per_gpu = val / world if batch_total is None else val * nrows_gpu / max(tot, 1)
            rec = {"tag": otag, "n": op["n"], "q": op["q"], "rows_per_gpu": nrows_gpu, "rows_total": tot, "value": val,
                   "value_per_gpu": per_gpu,
                   "unit": UNIT, "us_per_launch": o_ms * 1e3 / reps, "launches": reps,
                   "hbm_frac": per_gpu * 3 * op["n"] * pl.word_bytes / 1e9 / peaks["hbm_gbs"],
                   "kernel_variant": dict(pl.variants()).get(pl.default_variant, "?").split(" ")[0]}
            if peak_lo:
                rec["imad32_frac"] = per_gpu * MODMULS[op["n"]] * IMAD_PER_MODMUL[pl.word_bytes] / peak_lo
            tj = traffic_db.get(otag)
            if tj:
                rec["dram_bytes_per_row_ncu"] = tj["dram_bytes_per_row"]
            del x, y, z
            return rec

        for otag, nrows in OTHER_FIXED:
            others.append(one(otag, nrows))
        for B in SWEEP_BATCHES:      # the sweep's batch is the TOTAL batch: rank r takes its contiguous share
            per = (B + world - 1) // world
            mine = max(0, min(per, B - rank * per))
            rec = one(SWEEP_TAG, mine, batch_total=B)
            rec["sweep"] = "batch size (total over all GPUs)"
            others.append(rec)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- rooflines --------------------------------------------------------------------------------------------
    peaks, peak_src = measured_peaks()
    launch_s = ms * 1e-3 / args.steps
    per_gpu = value / world
    alg_bytes = 3 * n * wb * rows                                  # read a, read b, write c (per launch, per GPU)
    achieved = alg_bytes / launch_s / 1e9
    tj = load_json("traffic.json").get(tag)
    traffic = tj["dram_bytes_per_row"] * rows if tj else None
    roofline_hbm = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peak_src,
                    "bytes_per_polymul": 3 * n * wb, "launch_ms": launch_s * 1e3}
    # What binds is the integer multiplier, not HBM.  Its ceiling for THIS kernel follows from
    # the executed-instruction mix of the launched variant (profiles/sass_slots.json, written by tools/sass_slots.py
    # from an ncu report; histogram in profiles/r02_sass_hist_<variant>.txt): 4 pipe cycles per IMAD.WIDE / IMAD.HI
    # and 2 per other IMAD, on one 16-lane multiplier per SM sub-partition.
    sm_mhz = (clocks or {}).get("sm_max_mhz") or 1965.0
    sms = torch.cuda.get_device_properties(local).multi_processor_count
    slots = load_json("sass_slots.json").get(tag)
    if slots and slots.get("variant") == variant_name:
        cyc = slots["warps_per_row"] * slots["pipe_cycles_per_warp"]         # multiplier cycles one polymul occupies
        peak_cycles = sms * 4 * sm_mhz * 1e6
        roofline = {"bound": "int_mul_pipe", "achieved": per_gpu * cyc / 1e9, "peak": peak_cycles / 1e9,
                    "unit": "G multiplier-pipe cycles/s", "frac": per_gpu * cyc / peak_cycles, "traffic": traffic,
                    "kernel": variant_name, "launch_ms": launch_s * 1e3,
                    "ceiling_polymul_per_s": peak_cycles / cyc, "pipe_cycles_per_polymul": cyc,
                    "imad_wide_per_warp": slots["imad_wide_per_warp"], "imad_narrow_per_warp": slots["imad_narrow_per_warp"],
                    "inst_per_warp": slots["inst_per_warp"], "warps_per_polymul": slots["warps_per_row"],
                    "peak_source": f"{sms} SMs x 4 sub-partitions x {sm_mhz:.0f} MHz; slot counts from profiles/{slots['source']}"}
    else:
        roofline = dict(roofline_hbm, kernel=variant_name,
                        note="no executed-instruction histogram committed for this variant (tools/sass_slots.py): HBM roofline only")
    for rec in others or []:      # the same ceiling for the other configurations' kernels
        sl = load_json("sass_slots.json").get(rec["tag"])
        if sl and sl.get("variant") == rec["kernel_variant"]:
            rec["mul_pipe_frac"] = rec["value_per_gpu"] * sl["warps_per_row"] * sl["pipe_cycles_per_warp"] / (sms * 4 * sm_mhz * 1e6)
    # SURVEY section 8(d)'s accounting beside it: IMAD32 per polymul x polymul/s against the IMAD.LO rate measured in this run
    roofline_imad32 = None
    try:
        peak_lo, peak_wide = tntt.microbench(0, local), tntt.microbench(1, local)
        per_polymul = MODMULS[n] * IMAD_PER_MODMUL[wb]
        roofline_imad32 = {"bound": "imad32 (SURVEY 8d accounting)", "achieved": per_gpu * per_polymul / 1e12,
                           "peak": peak_lo / 1e12, "unit": "TIMAD32/s", "frac": per_gpu * per_polymul / peak_lo,
                           "imad32_per_polymul": per_polymul, "measured_imad_lo_per_s": peak_lo,
                           "measured_imad_wide_per_s": peak_wide,
                           "peak_source": "measured in this run (tntt_microbench, 8 CTAs x 256 threads per SM, ILP 8)"}
    except Exception as exc:  # measurement aid only
        roofline_imad32 = {"error": str(exc)}

    cpu = None
    os.sched_setaffinity(0, all_cpus)       # the CPU baseline gets every host core again
    if world == 1 and not args.no_cpu_baseline:
        ctx = cpu_reference_throughput(tag, seconds=10.0)
        t = time.perf_counter()
        ctx["run"](ctx["a"], ctx["b"])
        dt = time.perf_counter() - t
        cpu = {"value": ctx["rows"] / dt, "unit": UNIT, "cores": ctx["cores"], "kind": ctx["kind"], "sample_rows": ctx["rows"],
               "sample": f"{ctx['rows']} polymuls of the same workload, {ctx['name']}, one pass after warm-up"}
        # the Python golden model (new_reference/cg_ntt.py:78-92, restated in oracle/ntt_oracle.py), one core
        from oracle import ntt_oracle as O

        pa, pb = [int(v) for v in ctx["a"][0]], [int(v) for v in ctx["b"][0]]
        t = time.perf_counter()
        reps = 0
        while reps < 1 or (time.perf_counter() - t < 2.0 and reps < 50):
            O.nwc_poly_mult(pa, pb, psi, q)
            reps += 1
        cpu["python_reference"] = {"value": reps / (time.perf_counter() - t), "unit": UNIT, "cores": 1, "kind": "port",
                                   "sample": f"{reps} polymul(s), oracle/ntt_oracle.py nwc_poly_mult (pure Python)"}
        # the reference's C++ code for the other configurations, beside their GPU numbers (config 5: software_benchmark/
        # benchmark_ntt.cpp:279-284's timed loop via oracle/_ref)
        if others:
            done = {}
            for rec in others:
                if rec["tag"] not in done:
                    octx = cpu_reference_throughput(rec["tag"], seconds=2.0)
                    t = time.perf_counter()
                    octx["run"](octx["a"], octx["b"])
                    done[rec["tag"]] = {"value": octx["rows"] / (time.perf_counter() - t), "unit": UNIT, "cores": octx["cores"],
                                        "kind": octx["kind"], "sample_rows": octx["rows"], "implementation": octx["name"]}
                rec["cpu_reference"] = done[rec["tag"]]

    line = {
        "metric": METRIC if tag == "n4096_60" else f"polymuls_per_sec_{tag}", "value": value, "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64" if wb == 8 else "u32", "data": "synthetic",
        "config": workload_config(tag, rows, world),
        "run": {"parallelism": f"batch-sharded x{world}, no collective", "kernel_variant": variant_desc, "host_affinity": numa,
                "parity_gate": gate},
        "roofline": roofline, "roofline_hbm": roofline_hbm, "roofline_imad32": roofline_imad32, "sustained": sustained,
        "cpu_baseline": cpu, "e2e": e2e, "transform_domain": extras, "other_configs": others,
        "gpu_launches": launches, "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
