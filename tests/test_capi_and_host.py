"""CPU-only checks: the C-ABI library loads and exports every symbol include/tntt.h declares, the
host-side mirror raises the reference's errors, there is no CPU fallback, sharding logic."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "tntt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tntt_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import tntt

    assert os.path.exists(tntt.LIB_PATH), "build libtntt.so first (python -c 'import __graft_entry__ as g; g.build()')"
    L = ctypes.CDLL(tntt.LIB_PATH)
    declared = header_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/tntt.h but not exported"
    assert sorted(tntt.SYMBOLS) == declared, "python binding and header disagree"
    assert tntt.lib().tntt_version() == 200
    # nothing but the C ABI leaks out of the shared object
    out = subprocess.check_output(["nm", "-D", "--defined-only", tntt.LIB_PATH], text=True)
    exported = sorted(line.split()[-1] for line in out.splitlines() if " T " in line)
    assert [e for e in exported if e.startswith("tntt_")] == declared
    assert not [e for e in exported if not e.startswith("tntt_") and not e.startswith("_")]


def test_library_contains_sm100a_code_only():
    import tntt

    out = subprocess.run(["cuobjdump", "-lelf", tntt.LIB_PATH], capture_output=True, text=True).stdout
    if not out.strip():
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_a_gpu():
    import torch

    import cg_ntt
    import tntt

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cg_ntt.nwc_poly_mult([0] * 256, [0] * 256, 1239911)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cg_ntt.cg_ntt([0] * 256, 5)
    h = ctypes.c_void_p()
    rc = tntt.lib().tntt_plan_create(ctypes.byref(h), 0, 256, 8380417, 1239911, 1)
    assert rc == -7 and b"no CPU path" in tntt.lib().tntt_last_error()      # TNTT_NO_DEVICE
    v = ctypes.c_double()
    assert tntt.lib().tntt_microbench(0, 0, ctypes.byref(v)) == -7
    # the round-2 entry points fail the same way: no plan without a device, nothing computes on the host
    q, psi = (ctypes.c_uint64 * 1)(8380417), (ctypes.c_uint64 * 1)(1239911)
    rc = tntt.lib().tntt_rns_plan_create(ctypes.byref(h), 0, 256, q, psi, 1)
    assert rc == -7 and b"no CPU path" in tntt.lib().tntt_last_error()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tntt.RnsContext(256, [8380417], [1239911])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tntt.get_plans(256, 8380417, 1239911)
    assert tntt.lib().tntt_polymul_host_multi(None, 0, None, None, None, 0) == -1      # TNTT_BAD_ARG: no plans
    assert tntt.lib().tntt_polymul_spectrum_host(None, None, None, 1, None, 1) == -1   # TNTT_BAD_ARG: no plan


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tiny-ntt_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower().replace("no oracle", ""), f"{f} mentions the oracle"
                assert "/root/reference" not in text, f


def test_reference_error_behaviour_on_the_host_side():
    import cg_ntt
    import cg_ntt_8butterfly as m8

    assert (cg_ntt.N, cg_ntt.Q) == (256, 8380417)
    with pytest.raises(ValueError, match=r"^Expected 256 coefficients, got 3$"):
        cg_ntt.cg_ntt([1, 2, 3], 5)
    with pytest.raises(ValueError, match=r"^Expected 256 coefficients, got 0$"):
        cg_ntt.cg_intt([], 5)
    with pytest.raises(ValueError, match=r"^Expected 256 coefficients$"):
        cg_ntt.nwc_poly_mult([0] * 255, [0] * 256, 7)
    with pytest.raises(ValueError, match=r"^Expected 8 butterfly lanes$"):
        m8.butterfly_batch([0] * 8, [0] * 7, [0] * 8)
    with pytest.raises(ValueError, match=r"^Expected 256 coefficients, got 2$"):
        m8.cg_ntt_8butterfly([1, 2], 5)
    saved = cg_ntt.N
    try:
        cg_ntt.N = 1024                      # module globals are read at call time (cg_ntt.py:36)
        with pytest.raises(ValueError, match=r"^Expected 1024 coefficients, got 256$"):
            cg_ntt.cg_ntt([0] * 256, 5)
    finally:
        cg_ntt.N = saved
    assert cg_ntt.modinv(256) == 8347681 and cg_ntt.modinv(3, 7) == 5
    assert cg_ntt.bit_reverse(1, 8) == 128 and cg_ntt.bit_reverse(6, 3) == 3
    assert cg_ntt.bit_reverse_list(list(range(8))) == [0, 4, 2, 6, 1, 5, 3, 7]
    assert m8.N == 256 and m8.modinv is cg_ntt.modinv


def test_shard_ranges_cover_the_batch_exactly():
    from tntt.shard import shard_range

    for total in (0, 1, 7, 8, 9, 4096, 32768 + 5):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) == -(-total // world) if total else max(sizes) == 0
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tiny-ntt_b200"))
import numpy as np, torch, torch.distributed as dist
from tntt.shard import shard_range, shard_rows, max_over_ranks, sum_over_ranks
from oracle.cpu_ref import COracle
from oracle import ntt_oracle as O
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
p = O.PARAMS["dilithium"]; n, q, psi = p["n"], p["q"], p["psi"]
rng = np.random.default_rng(99)                      # same global batch on both ranks
a = rng.integers(0, q, size=(11, n), dtype=np.uint64); b = rng.integers(0, q, size=(11, n), dtype=np.uint64)
mine_a, mine_b = shard_rows(a, 2, rank), shard_rows(b, 2, rank)
lo, hi = shard_range(11, 2, rank)
assert mine_a.shape[0] == hi - lo == (6 if rank == 0 else 5)
# each rank computes ITS rows only (the oracle stands in for the GPU kernel here); no data-path collective
mine_c = COracle().nwc_poly_mult(mine_a, mine_b, psi, q)
full = COracle().nwc_poly_mult(a, b, psi, q)
assert (mine_c == full[lo:hi]).all()
# bench.py's timing reduction: max over ranks of the per-rank device time, sum of rows
assert max_over_ranks(1.0 + rank) == 2.0
assert sum_over_ranks(hi - lo) == 11.0
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_sharded_batches_with_gloo_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"rank {r} ok" in o


def test_benchmark_fixture_helpers_reproduce_the_reference_inputs_and_checksums(golden_all):
    # tntt.fixtures = make_poly / checksum of the C++ benchmarks, pinned by the golden
    # vectors that were produced by running the reference itself
    from tntt import fixtures

    for tag, g in golden_all.items():
        n, q = g["n"], g["q"]
        case = g["cases"]["lcg_1_2"]
        assert fixtures.make_poly(1, n, q) == case["a"], tag
        assert fixtures.make_poly(2, n, q) == case["b"], tag
        assert fixtures.checksum(case["c"], q) == fixtures.REFERENCE_CHECKSUMS[(n, q)], tag
        assert fixtures.REFERENCE_CHECKSUMS[(n, q)] in [int(v) for v in g["cpp_checksums"].values()], tag


def test_find_psi_matches_the_reference_script():
    """tntt_find_psi (host side) against answers of scripts/find_psi.py:9-44 run in the build container
    (tests/golden/make_find_psi_golden.py)."""
    import ctypes as C
    import json

    import tntt

    with open(os.path.join(ROOT, "tests", "golden", "golden_find_psi.json")) as fh:
        cases = json.load(fh)["cases"]
    L = tntt.lib()
    assert len(cases) >= 7
    for c in cases:
        psi = C.c_uint64()
        rc = L.tntt_find_psi(c["n"], c["q"], 10000, C.byref(psi))
        if c["psi"] is None:        # the script gives up; the library goes on and says so (rc = 1)
            assert rc == 1
            assert pow(psi.value, c["n"], c["q"]) == c["q"] - 1
        else:
            assert rc == 0 and psi.value == c["psi"], c
        assert tntt.find_psi(c["n"], c["q"]) == psi.value
    psi = C.c_uint64()
    assert L.tntt_find_psi(256, 7683, 10000, C.byref(psi)) < 0          # 7683 = 3 * 13 * 197
    assert L.tntt_find_psi(256, 7687, 10000, C.byref(psi)) < 0          # prime, but not 1 mod 512
    with pytest.raises(ValueError):
        tntt.find_psi(4096, 12289)


def test_bench_contract_of_the_reference_arm_and_shared_config():
    """bench.py --impl reference (CPU only) prints ONE JSON line with the contract's keys, and its `config` object is the
    one the GPU arm builds for the same workload (the driver compares the two)."""
    import json
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--config", "dilithium"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["sample_rows"] > 0
    sys.path.insert(0, ROOT)
    import bench

    assert d["config"] == bench.workload_config("dilithium", bench.ROWS["dilithium"], 1)
    assert set(d["config"]) == {"workload", "n", "q", "psi", "rows_per_gpu", "rows_total", "gpus", "l2"}
    # the sweep of BASELINE config 5 splits its TOTAL batch into contiguous shares, ranks beyond the batch idle
    for world in (1, 2, 4, 8):
        for B in bench.SWEEP_BATCHES:
            per = (B + world - 1) // world
            shares = [max(0, min(per, B - r * per)) for r in range(world)]
            assert sum(shares) == B and all(s >= 0 for s in shares)
