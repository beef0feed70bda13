#!/usr/bin/env python3
"""Launch every kernel of the library once on tiny inputs (for compute-sanitizer runs)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tiny-ntt_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import tntt  # noqa: E402
from bench import PARAMS  # noqa: E402
from oracle.cpu_ref import COracle  # noqa: E402

co = COracle()
launches = 0
for tag, p in PARAMS.items():
    plan = tntt.get_plan(p["n"], p["q"], p["psi"], True)
    rng = np.random.default_rng(1)
    rows = 19
    a = rng.integers(0, p["q"], size=(rows, p["n"]), dtype=np.uint64)
    b = rng.integers(0, p["q"], size=(rows, p["n"]), dtype=np.uint64)
    npdt, sdt = (np.uint32, np.int32) if plan.word_bytes == 4 else (np.uint64, np.int64)
    da = torch.from_numpy(a.astype(npdt).view(sdt)).cuda()
    db = torch.from_numpy(b.astype(npdt).view(sdt)).cuda()
    want = co.nwc_poly_mult(a, b, p["psi"], p["q"], threads=4)
    for vid, desc in plan.variants():
        got = tntt.polymul(plan, da, db, variant=vid).cpu().numpy().view(npdt).astype(np.uint64)
        assert (got == want).all(), desc
        launches += 1
    f = tntt.forward(plan, da, twist=True)
    assert torch.equal(tntt.inverse(plan, f, twist=True), da)
    tntt.pointwise(plan, da, db)
    cur = tntt.bit_reverse(plan, da)
    for s in range(1, plan.logn + 1):
        cur = tntt.cg_stage(plan, cur, s)
    assert torch.equal(cur, tntt.forward(plan, da))
    tntt.scale(plan, da, 12345)
    tntt.reduce(plan, da)
    launches += 6 + plan.logn
torch.cuda.synchronize()
print("sanitize_all ok:", launches, "launches")
