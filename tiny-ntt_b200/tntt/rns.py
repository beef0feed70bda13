"""Multi-modulus (RNS) batches: SURVEY.md section 8 row f3.

The reference's reports name RNS / FHE parameter sets as the next step after the single 60-bit modulus
(reports/final-report.tex:1811,1817).  Here a residue-number-system polynomial is a tensor ``[L, B, N]``:
limb ``l`` holds the coefficients mod ``q_l``.  Limbs are independent, so a product is one fused polymul
launch per limb on that limb's plan (all limbs must share a word size); nothing crosses limbs on the device.
CRT reconstruction to big integers is host-side (Python ints) and meant for tests and small results.
"""
from __future__ import annotations

from typing import List, Sequence

from . import ops
from .plan import Plan, get_plan


def find_psi(n: int, q: int, start: int = 2, limit: int = 1 << 20) -> int:
    """A primitive 2n-th root of unity mod q (scripts/find_psi.py:9-44, without the 10 000 search cap:
    candidates g^((q-1)/2n) are tried instead of testing every integer)."""
    if (q - 1) % (2 * n):
        raise ValueError(f"q = {q} is not 1 mod 2n = {2 * n}: no primitive 2n-th root exists")
    e = (q - 1) // (2 * n)
    for g in range(start, limit):
        psi = pow(g, e, q)
        if pow(psi, n, q) == q - 1:
            return psi
    raise ValueError("no primitive root found")


class RnsContext:
    """Plans for the ring Z_Q[x]/(x^n+1), Q = prod(q_l), one limb per prime q_l."""

    def __init__(self, n: int, moduli: Sequence[int], psis: Sequence[int] | None = None, device: int | None = None):
        if not moduli:
            raise ValueError("need at least one modulus")
        self.n = n
        self.moduli = [int(q) for q in moduli]
        self.psis = [int(p) for p in psis] if psis is not None else [find_psi(n, q) for q in self.moduli]
        self.plans: List[Plan] = [get_plan(n, q, p, True, device) for q, p in zip(self.moduli, self.psis)]
        words = {pl.word_bytes for pl in self.plans}
        if len(words) != 1:
            raise ValueError("all limbs must use the same word size (mix of 32- and 64-bit moduli)")
        self.word_bytes = words.pop()
        self.Q = 1
        for q in self.moduli:
            self.Q *= q

    @property
    def dtype(self):
        return self.plans[0].dtype

    def _check(self, t, name):
        if t.dim() != 3 or t.shape[0] != len(self.plans) or t.shape[-1] != self.n:
            raise ValueError(f"{name} must have shape [L={len(self.plans)}, B, N={self.n}], got {tuple(t.shape)}")

    def polymul(self, a, b, out=None):
        """Negacyclic product limb by limb: out[l] = a[l] * b[l] in Z_{q_l}[x]/(x^n+1)."""
        import torch

        ta, tb = ops.as_tensor(a), ops.as_tensor(b)
        self._check(ta, "a")
        self._check(tb, "b")
        o = torch.empty_like(ta) if out is None else out
        for l, plan in enumerate(self.plans):
            ops.polymul(plan, ta[l], tb[l], out=o[l])
        return o

    def forward(self, a, twist: bool = True, out=None):
        import torch

        ta = ops.as_tensor(a)
        self._check(ta, "a")
        o = torch.empty_like(ta) if out is None else out
        for l, plan in enumerate(self.plans):
            ops.forward(plan, ta[l], twist=twist, out=o[l])
        return o

    def inverse(self, a, twist: bool = True, out=None):
        import torch

        ta = ops.as_tensor(a)
        self._check(ta, "a")
        o = torch.empty_like(ta) if out is None else out
        for l, plan in enumerate(self.plans):
            ops.inverse(plan, ta[l], twist=twist, out=o[l])
        return o

    # ---- operands kept in the transform domain (one spectrum per limb; see ops.forward_spectrum) ----
    def _per_limb(self, fn, a, out, *rest):
        import torch

        ta = ops.as_tensor(a)
        self._check(ta, "a")
        o = torch.empty_like(ta) if out is None else out
        for l, plan in enumerate(self.plans):
            fn(plan, ta[l], *[r[l] for r in rest], out=o[l])
        return o

    def forward_spectrum(self, a, out=None):
        return self._per_limb(ops.forward_spectrum, a, out)

    def inverse_spectrum(self, a, out=None):
        return self._per_limb(ops.inverse_spectrum, a, out)

    def pointwise(self, a, b, out=None):
        tb = ops.as_tensor(b)
        self._check(tb, "b")
        return self._per_limb(ops.pointwise, a, out, tb)

    def polymul_spectrum(self, a, b_spectrum, out=None):
        """out[l] = a[l] * b[l] with b given as per-limb spectra, shape [L, B, N] or [L, 1, N] (shared by the batch)."""
        tb = ops.as_tensor(b_spectrum)
        if tb.dim() != 3 or tb.shape[0] != len(self.plans) or tb.shape[-1] != self.n:
            raise ValueError(f"b_spectrum must have shape [L={len(self.plans)}, B or 1, N={self.n}]")
        return self._per_limb(ops.polymul_spectrum, a, out, tb)

    # ---- host-side helpers (tests, small data) ------------------------------------------------
    def decompose(self, coeffs: Sequence[Sequence[int]]):
        """[B][N] big integers -> [L, B, N] residues (numpy, word dtype)."""
        import numpy as np

        dt = np.uint32 if self.word_bytes == 4 else np.uint64
        return np.array([[[int(c) % q for c in row] for row in coeffs] for q in self.moduli], dtype=dt)

    def reconstruct(self, residues) -> List[List[int]]:
        """[L, B, N] residues -> [B][N] integers in [0, Q) by the Chinese remainder theorem."""
        L = len(self.moduli)
        inv = [pow(self.Q // q, -1, q) * (self.Q // q) for q in self.moduli]
        res = [[[int(v) for v in row] for row in limb] for limb in residues]
        rows = len(res[0])
        return [[sum(res[l][r][i] * inv[l] for l in range(L)) % self.Q for i in range(self.n)] for r in range(rows)]
