import sys, os
ROOT="/root/repo"; sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tiny-ntt_b200"))
import torch, tntt
from bench import PARAMS
p = PARAMS["n4096_60"]; plan = tntt.get_plan(p["n"], p["q"], p["psi"], True)
rows = 16384
g = torch.Generator(device="cuda").manual_seed(1)
a = torch.randint(0, p["q"], (rows, p["n"]), generator=g, device="cuda", dtype=torch.int64)
b = torch.randint(0, p["q"], (rows, p["n"]), generator=g, device="cuda", dtype=torch.int64)
c = torch.empty_like(a); spec = tntt.forward_spectrum(plan, b)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return rows / (e0.elapsed_time(e1) / n * 1e-3) / 1e6
for rep in range(3):
    print("fwd_spec %.2f  inv_spec %.2f  pm_spec %.2f  pm_spec_shared %.2f  polymul %.2f" % (
        t(lambda: tntt.forward_spectrum(plan, a, out=c)), t(lambda: tntt.inverse_spectrum(plan, spec, out=c)),
        t(lambda: tntt.polymul_spectrum(plan, a, spec, out=c)), t(lambda: tntt.polymul_spectrum(plan, a, spec[0], out=c)),
        t(lambda: tntt.polymul(plan, a, b, out=c))))
