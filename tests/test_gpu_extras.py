"""GPU tests of the rows next to the hot path: the plain-C client, RNS batches, the TMA variant."""
import os
import subprocess

import numpy as np
import pytest

from oracle import ntt_oracle as O
from oracle.cpu_ref import COracle

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plain_c_client_follows_the_rocc_sequence(tmp_path):
    """examples/rocc_style_driver.c: load A, load B, start, read -- the vector of chipyard/ntt-test.c:101-107."""
    exe = str(tmp_path / "rocc_driver")
    lib = os.path.join(ROOT, "tiny-ntt_b200")
    subprocess.check_call(["gcc", "-O2", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "rocc_style_driver.c"),
                           "-L" + lib, "-ltntt", "-Wl,-rpath," + lib, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "c[0..3] = 5 11 17 3" in out.stdout and "PASS" in out.stdout


def test_rns_product_matches_bigint_schoolbook():
    import tntt

    n = 256
    # three NTT-friendly 60-bit primes q = k * 2^20 + 1 below 2^60 (found by trial, checked prime by the library)
    moduli = []
    k = (1 << 40) - 1
    while len(moduli) < 3:
        q = k * (1 << 20) + 1
        if q < (1 << 60) and all(q % p for p in (3, 5, 7, 11, 13)) and pow(2, q - 1, q) == 1:
            try:
                tntt.find_psi(n, q)
                moduli.append(q)
            except ValueError:
                pass
        k -= 1
    ctx = tntt.RnsContext(n, moduli)
    assert ctx.word_bytes == 8
    rng = np.random.default_rng(3)
    rows = 3
    big_a = [[int(rng.integers(0, 1 << 62)) * int(rng.integers(0, 1 << 62)) % ctx.Q for _ in range(n)] for _ in range(rows)]
    big_b = [[int(rng.integers(0, 1 << 62)) for _ in range(n)] for _ in range(rows)]
    a = torch.from_numpy(ctx.decompose(big_a).view(np.int64)).cuda()
    b = torch.from_numpy(ctx.decompose(big_b).view(np.int64)).cuda()
    c = ctx.polymul(a, b)
    got = ctx.reconstruct(c.cpu().numpy().view(np.uint64))
    for r in range(rows):
        assert got[r] == O.schoolbook_negacyclic(big_a[r], big_b[r], ctx.Q)
    # transform-domain round trip per limb
    assert torch.equal(ctx.inverse(ctx.forward(a)), a)
    # operands kept as spectra: same products
    sb = ctx.forward_spectrum(b)
    assert torch.equal(ctx.polymul_spectrum(a, sb), c)
    assert torch.equal(ctx.inverse_spectrum(ctx.pointwise(ctx.forward_spectrum(a), sb)), c)
    shared = ctx.polymul_spectrum(a, sb[:, :1])
    assert torch.equal(shared, ctx.polymul(a, b[:, :1].expand_as(b).contiguous()))


def test_tma_variant_agrees_and_is_listed():
    import tntt

    p = O.PARAMS["n4096_60"]
    plan = tntt.get_plan(p["n"], p["q"], p["psi"], True)
    tma = [v for v, d in plan.variants() if d.split(" ")[0].endswith("_t1")]
    assert tma, "no TMA-staged variant built"
    rng = np.random.default_rng(8)
    a = rng.integers(0, p["q"], size=(300, p["n"]), dtype=np.uint64)
    b = rng.integers(0, p["q"], size=(300, p["n"]), dtype=np.uint64)
    want = COracle().nwc_poly_mult(a, b, p["psi"], p["q"], threads=8)
    da, db = torch.from_numpy(a.view(np.int64)).cuda(), torch.from_numpy(b.view(np.int64)).cuda()
    for v in tma:
        for _ in range(3):   # back-to-back launches reuse the mbarrier phase logic from scratch each time
            got = tntt.polymul(plan, da, db, variant=v).cpu().numpy().view(np.uint64)
            assert (got == want).all()


def test_refs_twins_match_the_reference_outputs():
    # tests/golden/golden_refs.json was produced by the reference's own test/refs modules
    # (tests/golden/make_golden_refs.py); inputs are unreduced / negative on purpose
    import json

    from refs import ntt_forward_reference, ntt_inverse_reference
    from refs.ntt_forward_reference import bit_reverse_order

    with open(os.path.join(ROOT, "tests", "golden", "golden_refs.json")) as fh:
        cases = json.load(fh)
    for c in cases:
        n, q, psi = c["n"], c["q"], c["psi"]
        assert ntt_forward_reference(c["x"], n, q, psi) == c["forward"], (n, q)
        assert ntt_inverse_reference(c["x"], n, q, psi) == c["inverse"], (n, q)
        assert ntt_inverse_reference(c["forward"], n, q, psi) == [v % q for v in c["x"]]
    with pytest.raises(ValueError, match="Input must have 256 coefficients, got 3"):
        ntt_forward_reference([1, 2, 3], 256, 8380417, 1239911)
    with pytest.raises(ValueError, match="Input must have 4096 coefficients, got 2"):
        ntt_inverse_reference([1, 2])                      # module defaults: NTT_N = 4096
    assert bit_reverse_order(8) == [0, 4, 2, 6, 1, 5, 3, 7]


@pytest.mark.parametrize("tag", ["n4096_60", "n4096_24"])
def test_batch_size_dispatch_cluster_and_small_shapes(tag):
    # tntt_polymul picks the cluster kernel (one row per 4-CTA cluster, DSMEM exchanges) for tiny batches and
    # the one-CTA-per-SM shape up to one row per SM; every choice must give the oracle's bits
    import tntt

    p = O.PARAMS[tag]
    n, q, psi = p["n"], p["q"], p["psi"]
    tntt.clear_plan_cache()
    plan = tntt.get_plan(n, q, psi, True)
    names = dict((v, d.split()[0]) for v, d in plan.variants())
    assert any(nm.endswith("_c4") for nm in names.values())          # the cluster kernel exists for both word sizes
    if tag == "n4096_60":
        assert plan.cluster_variant >= 0 and names[plan.cluster_variant].endswith("_c4")
        assert plan.cluster_batch_max == torch.cuda.get_device_properties(0).multi_processor_count // 4
        assert plan.small_variant >= 0 and plan.small_batch_max >= plan.cluster_batch_max
    else:
        assert plan.cluster_variant == -1      # 32-bit rows: dispatch keeps the one-CTA kernel (measured no gain)
    co = COracle()
    npdt = np.uint32 if plan.word_bytes == 4 else np.uint64
    sdt = np.int32 if plan.word_bytes == 4 else np.int64
    rng = np.random.default_rng(11)
    edges = sorted({1, 2, max(plan.cluster_batch_max, 1), plan.cluster_batch_max + 1, max(plan.small_batch_max, 1),
                    plan.small_batch_max + 1, 150})
    for rows in edges:
        a = rng.integers(0, q, size=(rows, n), dtype=np.uint64)
        b = rng.integers(0, q, size=(rows, n), dtype=np.uint64)
        a[0], b[0] = q - 1, q - 1
        ta = torch.from_numpy(a.astype(npdt).view(sdt)).cuda()
        tb = torch.from_numpy(b.astype(npdt).view(sdt)).cuda()
        got = tntt.polymul(plan, ta, tb).cpu().numpy().view(npdt).astype(np.uint64)
        assert (got == co.nwc_poly_mult(a, b, psi, q, threads=8)).all(), rows
        # in place on the first operand
        tntt.polymul(plan, ta, tb, out=ta)
        assert (ta.cpu().numpy().view(npdt).astype(np.uint64) == got).all(), rows
    # an explicit default switches the dispatch off
    plan.set_default_variant(plan.default_variant)
    info = tntt._lib.PlanInfo()
    tntt.lib().tntt_plan_info_get(plan._h, __import__("ctypes").byref(info))
    assert info.cluster_variant == -1 and info.small_variant == -1
    tntt.clear_plan_cache()


@pytest.mark.parametrize("tag", ["dilithium", "n1024_24", "n4096_24", "n4096_60"])
def test_transform_domain_api(tag):
    # operands kept as spectra (tntt_spectrum_forward / _inverse / tntt_polymul_spectrum): every route to the
    # product gives the oracle's bits
    import tntt

    p = O.PARAMS[tag]
    n, q, psi = p["n"], p["q"], p["psi"]
    plan = tntt.get_plan(n, q, psi, True)
    assert plan.spectrum == 1
    co = COracle()
    npdt = np.uint32 if plan.word_bytes == 4 else np.uint64
    sdt = np.int32 if plan.word_bytes == 4 else np.int64
    rng = np.random.default_rng(5)
    rows = 37
    a = rng.integers(0, q, size=(rows, n), dtype=np.uint64)
    b = rng.integers(0, q, size=(rows, n), dtype=np.uint64)
    a[0], b[0] = q - 1, q - 1
    a[1], b[1] = O.make_poly(tag, 1), O.make_poly(tag, 2)
    dev = lambda v: torch.from_numpy(np.ascontiguousarray(v).astype(npdt).view(sdt)).cuda()      # noqa: E731
    host = lambda t: t.cpu().numpy().view(npdt).astype(np.uint64)                                # noqa: E731
    ta, tb = dev(a), dev(b)
    want = co.nwc_poly_mult(a, b, psi, q, threads=8)
    sa, sb = tntt.forward_spectrum(plan, ta), tntt.forward_spectrum(plan, tb)
    assert host(sa).max() < q
    assert sorted(host(sa)[1].tolist()) == sorted(host(tntt.forward(plan, ta[1:2], twist=True))[0].tolist())
    assert (host(tntt.inverse_spectrum(plan, sa)) == a).all()                                     # round trip
    assert (host(tntt.inverse_spectrum(plan, tntt.pointwise(plan, sa, sb))) == want).all()        # all in the transform domain
    assert (host(tntt.polymul_spectrum(plan, ta, sb)) == want).all()                              # one operand cached
    shared = co.nwc_poly_mult(a, np.broadcast_to(b[1], a.shape).copy(), psi, q, threads=8)
    assert (host(tntt.polymul_spectrum(plan, ta, sb[1])) == shared).all()                         # one spectrum for the batch
    assert (host(tntt.polymul_spectrum(plan, ta, sb[1:2])) == shared).all()
    assert host(tntt.polymul_spectrum(plan, ta[:1], sb[:1])).tolist() == want[:1].tolist()        # batch of one
    with pytest.raises(ValueError):
        tntt.polymul_spectrum(plan, ta, sb[:2])
    # plans without a fused size or without psi have no spectrum kernels: a loud error, no fallback
    small = tntt.get_plan(64, 8380417, pow(1239911, 4, 8380417), True)
    assert small.spectrum == 0
    with pytest.raises(tntt.TnttError):
        tntt.forward_spectrum(small, torch.zeros((1, 64), dtype=torch.int32, device="cuda"))


def test_entry_points_are_cuda_graph_capturable():
    # every fused-size entry point only enqueues kernels on the caller's stream (no allocation, no synchronisation),
    # so a launch-bound loop of small products can be captured once and replayed as a CUDA graph
    import tntt

    p = O.PARAMS["n4096_60"]
    n, q, psi = p["n"], p["q"], p["psi"]
    plan = tntt.get_plan(n, q, psi, True)
    co = COracle()
    rng = np.random.default_rng(9)
    rows = 4                                            # cluster-kernel territory
    a = rng.integers(0, q, size=(rows, n), dtype=np.uint64)
    b = rng.integers(0, q, size=(rows, n), dtype=np.uint64)
    ta = torch.from_numpy(a.view(np.int64)).cuda()
    tb = torch.from_numpy(b.view(np.int64)).cuda()
    c = torch.empty_like(ta)
    s = torch.empty_like(ta)
    r = torch.empty_like(ta)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):                       # warm-up outside capture, on the capture stream
        tntt.polymul(plan, ta, tb, out=c)
        tntt.forward_spectrum(plan, tb, out=s)
        tntt.polymul_spectrum(plan, ta, s, out=r)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    c.zero_()
    r.zero_()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        tntt.polymul(plan, ta, tb, out=c)
        tntt.forward_spectrum(plan, tb, out=s)
        tntt.polymul_spectrum(plan, ta, s, out=r)
        tntt.inverse(plan, tntt.forward(plan, ta), out=s)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    want = co.nwc_poly_mult(a, b, psi, q, threads=4)
    assert (c.cpu().numpy().view(np.uint64) == want).all()
    assert (r.cpu().numpy().view(np.uint64) == want).all()
    assert torch.equal(s, ta)
    # new inputs, same graph
    ta.copy_(tb)
    g.replay()
    torch.cuda.synchronize()
    assert (c.cpu().numpy().view(np.uint64) == co.nwc_poly_mult(b, b, psi, q, threads=4)).all()


Q60 = (1 << 60) - (1 << 14) + 1
OTHER_RINGS = [(512, 8380417, 1718063), (2048, 8380417, 7901702), (8192, 67043329, 8157893),
               (512, Q60, 984081769261068913), (2048, Q60, 644283108363935541), (8192, Q60, 527760526715669589)]


@pytest.mark.parametrize("n,q,psi", OTHER_RINGS)
def test_fused_kernels_for_other_ring_sizes(n, q, psi):
    # N = 512, 2048, 8192 (the sizes between and above the reference's three) have fused, transform-domain and
    # natural-order kernels too; everything against the oracle
    import tntt

    plan = tntt.get_plan(n, q, psi, True)
    assert plan.fused == 1 and plan.spectrum == 1
    co = COracle()
    omega = psi * psi % q
    npdt = np.uint32 if plan.word_bytes == 4 else np.uint64
    sdt = np.int32 if plan.word_bytes == 4 else np.int64
    rng = np.random.default_rng(n)
    rows = 19
    a = rng.integers(0, q, size=(rows, n), dtype=np.uint64)
    b = rng.integers(0, q, size=(rows, n), dtype=np.uint64)
    a[0], b[0] = q - 1, q - 1
    dev = lambda v: torch.from_numpy(np.ascontiguousarray(v).astype(npdt).view(sdt)).cuda()      # noqa: E731
    host = lambda t: t.cpu().numpy().view(npdt).astype(np.uint64)                                # noqa: E731
    ta, tb = dev(a), dev(b)
    want = co.nwc_poly_mult(a, b, psi, q, threads=8)
    for v, desc in plan.variants():
        assert (host(tntt.polymul(plan, ta, tb, variant=v)) == want).all(), desc
    assert (host(tntt.polymul(plan, ta, tb)) == want).all()
    sb = tntt.forward_spectrum(plan, tb)
    assert (host(tntt.polymul_spectrum(plan, ta, sb)) == want).all()
    assert (host(tntt.inverse_spectrum(plan, tntt.pointwise(plan, tntt.forward_spectrum(plan, ta), sb))) == want).all()
    assert (host(tntt.forward(plan, ta)) == co.cg_ntt(a, omega, q)).all()
    assert (host(tntt.inverse(plan, ta)) == co.cg_intt(a, omega, q)).all()
    assert torch.equal(tntt.inverse(plan, tntt.forward(plan, ta, twist=True), twist=True), ta)


@pytest.mark.parametrize("wb,logn,red", [(4, 8, 0), (4, 12, 0), (4, 13, 0), (8, 8, 0), (8, 12, 0), (8, 12, 1), (8, 13, 1)])
def test_largest_modulus_each_kernel_path_accepts(wb, logn, red):
    # tightest lazy ranges: the largest NTT-friendly prime a path admits, rows of all q-1; every variant, the
    # transform-domain kernels and the natural-order transforms against the oracle
    import emu
    import tntt

    n = 1 << logn
    q = emu.largest_modulus_of_path(wb, logn, red)
    psi = tntt.find_psi(n, q)
    plan = tntt.get_plan(n, q, psi, True)
    assert plan.word_bytes == wb and plan.lazy_reduce == red, (q, plan.word_bytes, plan.lazy_reduce)
    co = COracle()
    npdt = np.uint32 if wb == 4 else np.uint64
    sdt = np.int32 if wb == 4 else np.int64
    rng = np.random.default_rng(logn)
    rows = 21
    a = rng.integers(0, q, size=(rows, n), dtype=np.uint64)
    b = rng.integers(0, q, size=(rows, n), dtype=np.uint64)
    a[0], b[0] = q - 1, q - 1
    a[1] = q - 1
    ta = torch.from_numpy(a.astype(npdt).view(sdt)).cuda()
    tb = torch.from_numpy(b.astype(npdt).view(sdt)).cuda()
    host = lambda t: t.cpu().numpy().view(npdt).astype(np.uint64)      # noqa: E731
    want = co.nwc_poly_mult(a, b, psi, q, threads=8)
    for v, desc in plan.variants():
        assert (host(tntt.polymul(plan, ta, tb, variant=v)) == want).all(), (q, desc)
    assert (host(tntt.polymul_spectrum(plan, ta, tntt.forward_spectrum(plan, tb))) == want).all(), q
    assert (host(tntt.inverse(plan, ta)) == co.cg_intt(a, psi * psi % q, q)).all(), q
    assert torch.equal(tntt.inverse(plan, tntt.forward(plan, ta, twist=True), twist=True), ta)


@pytest.mark.parametrize("wb,logn,red", [(4, 14, 0), (4, 15, 0), (8, 14, 0), (8, 14, 1), (8, 15, 0), (8, 15, 1)])
def test_rows_longer_than_one_cta_run_on_clusters(wb, logn, red):
    # N = 16384 / 32768: a row does not fit one CTA's shared memory; the fused product runs one row per
    # thread-block cluster (4 / 8 CTAs, every exchange through distributed shared memory).  Largest admissible
    # modulus of each path, rows of all q-1, ragged batch sizes.
    import emu
    import tntt

    n = 1 << logn
    q = emu.largest_modulus_of_path(wb, logn, red)
    psi = tntt.find_psi(n, q)
    plan = tntt.get_plan(n, q, psi, True)
    assert plan.word_bytes == wb and plan.lazy_reduce == red and plan.fused == 1
    names = dict((v, d.split()[0]) for v, d in plan.variants())
    assert names[plan.default_variant].endswith("_c4" if logn == 14 else "_c8")
    co = COracle()
    npdt = np.uint32 if wb == 4 else np.uint64
    sdt = np.int32 if wb == 4 else np.int64
    rng = np.random.default_rng(logn)
    for rows in (1, 7, 40):
        a = rng.integers(0, q, size=(rows, n), dtype=np.uint64)
        b = rng.integers(0, q, size=(rows, n), dtype=np.uint64)
        a[0], b[0] = q - 1, q - 1
        ta = torch.from_numpy(a.astype(npdt).view(sdt)).cuda()
        tb = torch.from_numpy(b.astype(npdt).view(sdt)).cuda()
        got = tntt.polymul(plan, ta, tb).cpu().numpy().view(npdt).astype(np.uint64)
        want = co.nwc_poly_mult(a, b, psi, q, threads=8)
        assert (got == want).all(), (q, rows)
        # transform-domain kernels on clusters
        host = lambda t: t.cpu().numpy().view(npdt).astype(np.uint64)      # noqa: E731
        assert plan.spectrum == 1
        sa, sb = tntt.forward_spectrum(plan, ta), tntt.forward_spectrum(plan, tb)
        assert host(sa).max() < q
        assert torch.equal(tntt.inverse_spectrum(plan, sa), ta), (q, rows)
        assert (host(tntt.polymul_spectrum(plan, ta, sb)) == want).all(), (q, rows)
        assert (host(tntt.inverse_spectrum(plan, tntt.pointwise(plan, sa, sb))) == want).all(), (q, rows)
        if rows > 1:
            shared = co.nwc_poly_mult(a, np.broadcast_to(b[1], a.shape).copy(), psi, q, threads=8)
            assert (host(tntt.polymul_spectrum(plan, ta, sb[1])) == shared).all(), (q, rows)
    # natural-order transforms of these sizes run the literal schedule
    omega = psi * psi % q
    assert (host(tntt.forward(plan, ta[:2])) == co.cg_ntt(a[:2], omega, q)).all()
    assert sorted(host(sa)[0].tolist()) == sorted(host(tntt.forward(plan, ta[:1], twist=True))[0].tolist())


def test_plain_c_client_of_the_transform_domain_entry_points(tmp_path):
    """examples/cached_operand_driver.c: one key transformed once, many products through tntt_polymul_spectrum."""
    exe = str(tmp_path / "cached_operand")
    lib = os.path.join(ROOT, "tiny-ntt_b200")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    subprocess.check_call(["gcc", "-O2", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(cuda, "include"),
                           os.path.join(ROOT, "examples", "cached_operand_driver.c"), "-L" + lib, "-ltntt",
                           "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-Wl,-rpath," + lib, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "PASS" in out.stdout


def test_plain_c_client_of_the_multi_modulus_and_multi_gpu_entry_points(tmp_path):
    """examples/rns_driver.c: tntt_find_psi, tntt_rns_plan_create (device-generated tables), tntt_rns_polymul,
    tntt_rns_spectrum_forward + tntt_rns_polymul_spectrum, tntt_polymul_host_multi -- from plain C."""
    exe = str(tmp_path / "rns_driver")
    lib = os.path.join(ROOT, "tiny-ntt_b200")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    subprocess.check_call(["gcc", "-O2", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(cuda, "include"),
                           os.path.join(ROOT, "examples", "rns_driver.c"), "-L" + lib, "-ltntt",
                           "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-Wl,-rpath," + lib, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "PASS" in out.stdout and "same bits" in out.stdout


def test_benchmark_cli_prints_the_reference_report(tmp_path, golden_all):
    """examples/benchmark_ntt_gpu.c: the report of software_benchmark/benchmark_ntt.cpp:286-293 (and the 60-bit build) --
    same keys in the same order, and the two checksum lines the reference binaries print for each of the four rings,
    for one row and for a batch; --check compares with the schoolbook product like the reference's."""
    exe = str(tmp_path / "benchmark_ntt_gpu")
    lib = os.path.join(ROOT, "tiny-ntt_b200")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    subprocess.check_call(["gcc", "-O2", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(cuda, "include"),
                           os.path.join(ROOT, "examples", "benchmark_ntt_gpu.c"), "-L" + lib, "-ltntt",
                           "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-Wl,-rpath," + lib, "-o", exe])
    keys = ["forward_ntt_total_ns", "forward_ntt_avg_ns", "forward_ntt_checksum", "total_ns", "avg_ns", "checksum"]
    for tag, g in golden_all.items():
        for batch in (1, 37):
            args = [exe, "--check", "--reps", "3", "--n", str(g["n"]), "--q", str(g["q"]), "--psi", str(g["psi"]), "--batch", str(batch)]
            out = subprocess.run(args, capture_output=True, text=True, timeout=300)
            assert out.returncode == 0, (tag, out.stdout + out.stderr)
            lines = out.stdout.strip().splitlines()
            assert lines[0] == "benchmark_ntt_gpu" and lines[1] == f"N={g['n']} Q={g['q']} reps=3", (tag, lines[:2])
            assert [l.split("=")[0] for l in lines[2:8]] == keys, (tag, lines)
            report = dict(l.split("=") for l in lines[2:])
            assert int(report["checksum"]) == g["cpp_checksums"]["checksum"], tag
            assert int(report["forward_ntt_checksum"]) == g["cpp_checksums"]["forward_ntt_checksum"], tag
            assert (len(lines) == 8) == (batch == 1)
    # the reference's defaults (N=256, Q=8380417) and its usage error
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "N=256 Q=8380417 reps=100" in out.stdout
    assert f"checksum={golden_all['dilithium']['cpp_checksums']['checksum']}" in out.stdout.splitlines()
    bad = subprocess.run([exe, "--frobnicate"], capture_output=True, text=True, timeout=120)
    assert bad.returncode == 2 and bad.stderr.startswith("usage: benchmark [--check] [--reps count]")


# ------------------------------------------------------------------------------------------------
# multi-modulus engine (csrc/rns.cu): one launch for [L, B, N], tables generated on the device
# ------------------------------------------------------------------------------------------------
def friendly_primes(n, count, below):
    """`count` primes q = 1 mod 2n, descending from `below` (deterministic Miller-Rabin)."""
    def is_prime(m):
        if m < 2:
            return False
        for p in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
            if m % p == 0:
                return m == p
        d, r = m - 1, 0
        while d % 2 == 0:
            d, r = d // 2, r + 1
        for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
            x = pow(a, d, m)
            if x in (1, m - 1):
                continue
            for _ in range(r - 1):
                x = x * x % m
                if x == m - 1:
                    break
            else:
                return False
        return True

    out, q = [], below - (below - 1) % (2 * n)
    while len(out) < count:
        if q < below and is_prime(q):
            out.append(q)
        q -= 2 * n
    return out


@pytest.mark.parametrize("n,limbs,below", [(4096, 1, 1 << 60), (4096, 4, 1 << 60), (4096, 16, 1 << 60), (4096, 20, 1 << 59),
                                           (4096, 3, 1 << 50), (1024, 5, 1 << 60), (256, 4, 1 << 55),
                                           (4096, 6, 1 << 23), (1024, 18, 1 << 23), (256, 3, 1 << 23)])
def test_rns_one_launch_matches_oracle_limb_by_limb(n, limbs, below):
    import tntt

    co = COracle()
    moduli = friendly_primes(n, limbs, below)
    psis = [tntt.find_psi(n, q) for q in moduli]
    ctx = tntt.RnsContext(n, moduli, psis)
    assert ctx.word_bytes == (4 if below <= (1 << 23) else 8)
    for l in {0, limbs // 2, limbs - 1}:
        assert ctx.tables_match_host_generators(l), tntt.lib().tntt_last_error()
    rng = np.random.default_rng(n + limbs)
    npdt, sdt = (np.uint32, np.int32) if ctx.word_bytes == 4 else (np.uint64, np.int64)
    for rows in (1, 7, 33):
        a = np.stack([rng.integers(0, q, size=(rows, n), dtype=np.uint64) for q in moduli])
        b = np.stack([rng.integers(0, q, size=(rows, n), dtype=np.uint64) for q in moduli])
        for l, q in enumerate(moduli):
            a[l, 0], b[l, 0] = q - 1, q - 1                  # largest canonical values
        da = torch.from_numpy(a.astype(npdt).view(sdt)).cuda()
        db = torch.from_numpy(b.astype(npdt).view(sdt)).cuda()
        got = ctx.polymul(da, db).cpu().numpy().view(npdt).astype(np.uint64)
        for l, (q, psi) in enumerate(zip(moduli, psis)):
            assert (got[l] == co.nwc_poly_mult(a[l], b[l], psi, q, threads=8)).all(), (l, q)
    # the single-modulus plans (host-built tables) give the same bits
    c = ctx.polymul(da, db)
    single = torch.stack([tntt.polymul(pl, da[l], db[l]) for l, pl in enumerate(ctx.plans)])
    assert torch.equal(single, c)
    # transform-domain entry points, every limb per launch: same spectra as the single-modulus plans, same products
    sb = ctx.forward_spectrum(db)
    assert torch.equal(sb, torch.stack([tntt.forward_spectrum(pl, db[l]) for l, pl in enumerate(ctx.plans)]))
    assert torch.equal(ctx.inverse_spectrum(sb), db)
    assert torch.equal(ctx.polymul_spectrum(da, sb), c)
    assert torch.equal(ctx.inverse_spectrum(ctx.pointwise(ctx.forward_spectrum(da), sb)), c)
    shared = ctx.polymul_spectrum(da, sb[:, :1].contiguous())
    assert torch.equal(shared, ctx.polymul(da, db[:, :1].expand_as(db).contiguous()))


def test_rns_context_rejects_what_it_cannot_run():
    import tntt

    q60 = friendly_primes(4096, 2, 1 << 60)
    q23 = friendly_primes(4096, 1, 1 << 23)
    with pytest.raises(ValueError, match="word"):
        tntt.RnsContext(4096, q60 + q23)                       # mixed word sizes
    with pytest.raises(ValueError, match="repeats"):
        tntt.RnsContext(4096, [q60[0], q60[0]])
    with pytest.raises(ValueError, match="psi"):
        tntt.RnsContext(4096, q60, [3, 5])
    with pytest.raises(ValueError):
        tntt.RnsContext(4096, [q60[0] + 2])                    # not prime / not 1 mod 2n
    ctx = tntt.RnsContext(4096, q60)
    a = torch.zeros((2, 3, 4096), dtype=torch.int64, device="cuda")
    with pytest.raises(ValueError):
        ctx.polymul(a[:1], a[:1])
    with pytest.raises(TypeError):
        ctx.polymul(a.to(torch.int32), a.to(torch.int32))
    with pytest.raises(ValueError):
        ctx.polymul(a.cpu(), a.cpu())
    assert ctx.polymul(a[:, :0], a[:, :0]).shape == (2, 0, 4096)


def test_polymul_sharded_over_all_visible_gpus():
    """tntt_polymul_host_multi: one process, one plan per device, contiguous row ranges (ragged on purpose), host barrier.
    On a one-GPU box this is the single-plan path; with more devices visible every GPU takes its share."""
    import tntt

    p = O.PARAMS["n4096_60"]
    n, q, psi = p["n"], p["q"], p["psi"]
    plans = tntt.get_plans(n, q, psi)
    assert len(plans) == torch.cuda.device_count() and len({pl.device for pl in plans}) == len(plans)
    rng = np.random.default_rng(len(plans))
    rows = 37 * len(plans) + 5
    a = rng.integers(0, q, size=(rows, n), dtype=np.uint64)
    b = rng.integers(0, q, size=(rows, n), dtype=np.uint64)
    ha, hb = torch.from_numpy(a.view(np.int64)).pin_memory(), torch.from_numpy(b.view(np.int64)).pin_memory()
    hc = tntt.polymul_sharded(plans, ha, hb)
    assert (hc.numpy().view(np.uint64) == COracle().nwc_poly_mult(a, b, psi, q, threads=8)).all()
    out = torch.empty_like(ha)
    assert tntt.polymul_sharded(plans, ha[:3], hb[:3], out=out[:3]) is not None       # fewer rows than devices is fine
    assert torch.equal(out[:3], hc[:3])
    with pytest.raises(tntt.TnttError):
        tntt.polymul_sharded([plans[0], plans[0]], ha, hb)                             # one plan per device
    with pytest.raises(ValueError):
        tntt.polymul_sharded(plans, ha.cuda(), hb.cuda())
    with pytest.raises(ValueError):
        tntt.polymul_sharded(plans, ha, hb, out=torch.empty((1, n), dtype=torch.int64))
    other = tntt.get_plan(256, 8380417, 1239911, True)
    with pytest.raises((tntt.TnttError, ValueError, TypeError)):
        tntt.polymul_sharded([other], ha, hb)
