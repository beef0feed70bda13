"""Multi-modulus (RNS) batches: SURVEY.md section 8 row f3.

The reference's reports name RNS / FHE parameter sets as the next step after the single 60-bit modulus
(reports/final-report.tex:1811,1817).  Here a residue-number-system polynomial is a tensor ``[L, B, N]``:
limb ``l`` holds the coefficients mod ``q_l``.  The product of all limbs is ONE fused kernel launch per 16
limbs (``tntt_rns_polymul``: the limb index is a grid dimension, tables and modulus constants come out of the
kernel parameters), and the twiddle / Shoup tables of every limb are generated on the device by one launch at
plan creation (``csrc/rns.cu``).  The per-limb single-modulus plans behind ``forward`` / ``inverse`` /
``*_spectrum`` are created lazily, only if those calls are used.  CRT reconstruction to big integers is
host-side (Python ints) and meant for tests and small results.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

from . import ops
from ._lib import check, lib
from .plan import Plan, _require_cuda, get_plan

FIND_PSI_MAX_SEARCH = 10000      # scripts/find_psi.py:9


def find_psi(n: int, q: int, max_search: int = FIND_PSI_MAX_SEARCH) -> int:
    """A primitive 2n-th root of unity mod q: the smallest psi in [2, max_search) with psi^n = -1, which is
    what scripts/find_psi.py:9-44 returns; where the script gives up (none that small) the search goes on
    over g^((q-1)/2n).  Host-side (tntt_find_psi), like the script."""
    psi = C.c_uint64()
    rc = lib().tntt_find_psi(n, q, max_search, C.byref(psi))
    if rc < 0:
        raise ValueError(lib().tntt_last_error().decode(errors="replace"))
    return psi.value


class RnsContext:
    """The ring Z_Q[x]/(x^n+1), Q = prod(q_l), one limb per prime q_l; all limbs share a word size."""

    def __init__(self, n: int, moduli: Sequence[int], psis: Sequence[int] | None = None, device: int | None = None):
        if not moduli:
            raise ValueError("need at least one modulus")
        torch = _require_cuda()
        self.n = n
        self.moduli = [int(q) for q in moduli]
        self.psis = [int(p) for p in psis] if psis is not None else [find_psi(n, q) for q in self.moduli]
        if len(self.psis) != len(self.moduli):
            raise ValueError("one psi per modulus")
        self.device = torch.cuda.current_device() if device is None else int(device)
        L = len(self.moduli)
        qa, pa = (C.c_uint64 * L)(*self.moduli), (C.c_uint64 * L)(*self.psis)
        h = C.c_void_p()
        rc = lib().tntt_rns_plan_create(C.byref(h), self.device, n, qa, pa, L)
        if rc:
            raise ValueError(lib().tntt_last_error().decode(errors="replace"))
        self._h = h
        self.word_bytes = lib().tntt_rns_plan_word_bytes(h)
        self.kernel = lib().tntt_rns_plan_kernel(h).decode()
        self.table_bytes = lib().tntt_rns_plan_table_bytes(h)
        self._plans: List[Plan] | None = None
        self.Q = 1
        for q in self.moduli:
            self.Q *= q

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                lib().tntt_rns_plan_destroy(h)
            except Exception:
                pass

    @property
    def plans(self) -> List[Plan]:
        """Single-modulus plans, one per limb (host-built tables): only the per-limb transform calls need them."""
        if self._plans is None:
            self._plans = [get_plan(self.n, q, p, True, self.device) for q, p in zip(self.moduli, self.psis)]
        return self._plans

    def tables_match_host_generators(self, limb: int) -> bool:
        """Test hook (tntt_rns_plan_check_tables): device-generated tables == the host generators', word for word."""
        return lib().tntt_rns_plan_check_tables(self._h, limb) == 0

    def kernel_attributes(self):
        regs, local, occ = C.c_int(), C.c_size_t(), C.c_int()
        check(lib().tntt_rns_kernel_attributes(self._h, C.byref(regs), C.byref(local), C.byref(occ)))
        return {"kernel": self.kernel, "regs": regs.value, "local_bytes": local.value, "ctas_per_sm": occ.value}

    @property
    def dtype(self):
        import torch

        return torch.int32 if self.word_bytes == 4 else torch.int64

    def _check(self, t, name):
        import torch

        if t.dim() != 3 or t.shape[0] != len(self.moduli) or t.shape[-1] != self.n:
            raise ValueError(f"{name} must have shape [L={len(self.moduli)}, B, N={self.n}], got {tuple(t.shape)}")
        ok = (torch.int32, torch.uint32) if self.word_bytes == 4 else (torch.int64, torch.uint64)
        if t.dtype not in ok:
            raise TypeError(f"{name} has dtype {t.dtype}; this context uses {ok}")
        if not t.is_cuda or t.device.index != self.device:
            raise ValueError(f"{name} must live on cuda:{self.device} (no CPU path)")

    def polymul(self, a, b, out=None):
        """Negacyclic product of every limb: out[l] = a[l] * b[l] in Z_{q_l}[x]/(x^n+1); one launch per 16 limbs."""
        import torch

        ta, tb = ops.as_tensor(a), ops.as_tensor(b)
        self._check(ta, "a")
        self._check(tb, "b")
        if ta.shape != tb.shape or ta.dtype != tb.dtype:
            raise ValueError("a and b must have the same shape and dtype")
        ta, tb = ta.contiguous(), tb.contiguous()
        o = ops._out_like(ta, out)
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        check(lib().tntt_rns_polymul(self._h, ta.data_ptr(), tb.data_ptr(), o.data_ptr(), ta.shape[1], stream))
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

    # ---- operands kept in the transform domain: [L, B, N] spectra, every limb in one launch (csrc/rns.cu) ----
    def _stream(self):
        import torch

        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _unary(self, fn, a, out):
        ta = ops.as_tensor(a)
        self._check(ta, "a")
        ta = ta.contiguous()
        o = ops._out_like(ta, out)
        check(fn(self._h, ta.data_ptr(), o.data_ptr(), ta.shape[1], self._stream()))
        return o

    def forward_spectrum(self, a, out=None):
        """Coefficients -> spectra (tntt_rns_spectrum_forward): ntt(twist(.)) of every row of every limb, canonical, in the
        spectrum order of the single-modulus plans of the same size."""
        return self._unary(lib().tntt_rns_spectrum_forward, a, out)

    def inverse_spectrum(self, a, out=None):
        """Spectra -> coefficients (tntt_rns_spectrum_inverse)."""
        return self._unary(lib().tntt_rns_spectrum_inverse, a, out)

    def pointwise(self, a, b, out=None):
        """Element-wise product mod q_l of two [L, B, N] tensors (spectra or not): tntt_rns_pointwise."""
        ta, tb = ops.as_tensor(a), ops.as_tensor(b)
        self._check(ta, "a")
        self._check(tb, "b")
        if ta.shape != tb.shape or ta.dtype != tb.dtype:
            raise ValueError("a and b must have the same shape and dtype")
        ta, tb = ta.contiguous(), tb.contiguous()
        o = ops._out_like(ta, out)
        check(lib().tntt_rns_pointwise(self._h, ta.data_ptr(), tb.data_ptr(), o.data_ptr(), ta.shape[1], self._stream()))
        return o

    def polymul_spectrum(self, a, b_spectrum, out=None):
        """out[l] = a[l] * b[l] with b given as per-limb spectra, shape [L, B, N] or [L, 1, N] (shared by the batch)."""
        ta, tb = ops.as_tensor(a), ops.as_tensor(b_spectrum)
        self._check(ta, "a")
        self._check(tb, "b_spectrum")
        if tb.dtype != ta.dtype or tb.shape[1] not in (1, ta.shape[1]):
            raise ValueError(f"b_spectrum must have shape [L={len(self.moduli)}, B or 1, N={self.n}] and a's dtype")
        ta, tb = ta.contiguous(), tb.contiguous()
        o = ops._out_like(ta, out)
        check(lib().tntt_rns_polymul_spectrum(self._h, ta.data_ptr(), tb.data_ptr(), o.data_ptr(), ta.shape[1], tb.shape[1],
                                              self._stream()))
        return o

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
