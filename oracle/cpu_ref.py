"""ctypes loaders for the two compiled CPU checkers.  TEST INFRASTRUCTURE ONLY.

* ``COracle``  -- oracle/_build/liboracle.so, our plain-C restatement (ntt_oracle.c).
* ``RefLib``   -- oracle/_ref/libref_<tag>_<simd>.so, the UNMODIFIED reference
  benchmark sources compiled in place by oracle/Makefile (ref_shim.cpp).

Nothing under tiny-ntt_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

from .ntt_oracle import PARAMS

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")

_u64p = C.POINTER(C.c_uint64)


def build_oracle(force: bool = False) -> str:
    """Compile ntt_oracle.c (gcc).  Returns the .so path."""
    src = os.path.join(HERE, "ntt_oracle.c")
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    return ORACLE_SO


def build_ref(reference: str = "/root/reference") -> bool:
    """Compile the reference's own benchmark sources into oracle/_ref (needs the reference tree)."""
    if not os.path.isdir(os.path.join(reference, "software_benchmark")):
        return False
    subprocess.check_call(["make", "-s", "-j8", "-C", HERE, "ref", f"REFERENCE={reference}"])
    return True


def _as_u64(x) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x, dtype=np.uint64))


class COracle:
    """The plain-C restatement; all arrays are uint64."""

    def __init__(self) -> None:
        self.lib = C.CDLL(build_oracle())
        L = self.lib
        L.tntt_oracle_cg_ntt.argtypes = [_u64p, _u64p, C.c_uint32, C.c_uint64, C.c_uint64]
        L.tntt_oracle_cg_intt.argtypes = [_u64p, _u64p, C.c_uint32, C.c_uint64, C.c_uint64]
        L.tntt_oracle_nwc_poly_mult.argtypes = [_u64p, _u64p, _u64p, C.c_uint32, C.c_uint64, C.c_uint64]
        L.tntt_oracle_nwc_poly_mult_batch.argtypes = [_u64p, _u64p, _u64p, C.c_size_t, C.c_uint32, C.c_uint64,
                                                      C.c_uint64, C.c_int]
        L.tntt_oracle_schoolbook.argtypes = [_u64p, _u64p, _u64p, C.c_uint32, C.c_uint64]
        L.tntt_oracle_make_poly.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_int, _u64p]
        L.tntt_oracle_checksum.argtypes = [_u64p, C.c_uint32, C.c_int]
        L.tntt_oracle_checksum.restype = C.c_uint64

    @staticmethod
    def _p(a: np.ndarray):
        return a.ctypes.data_as(_u64p)

    def cg_ntt(self, a, omega: int, q: int) -> np.ndarray:
        a = _as_u64(a)
        out = np.empty_like(a)
        for r_in, r_out in zip(a.reshape(-1, a.shape[-1]), out.reshape(-1, a.shape[-1])):
            self.lib.tntt_oracle_cg_ntt(self._p(r_in), self._p(r_out), a.shape[-1], omega, q)
        return out

    def cg_intt(self, a, omega: int, q: int) -> np.ndarray:
        a = _as_u64(a)
        out = np.empty_like(a)
        for r_in, r_out in zip(a.reshape(-1, a.shape[-1]), out.reshape(-1, a.shape[-1])):
            self.lib.tntt_oracle_cg_intt(self._p(r_in), self._p(r_out), a.shape[-1], omega, q)
        return out

    def nwc_poly_mult(self, a, b, psi: int, q: int, threads: int = 1) -> np.ndarray:
        a, b = _as_u64(a), _as_u64(b)
        n = a.shape[-1]
        out = np.empty_like(a)
        self.lib.tntt_oracle_nwc_poly_mult_batch(self._p(a), self._p(b), self._p(out), a.size // n, n, psi, q, threads)
        return out

    def schoolbook(self, a, b, q: int) -> np.ndarray:
        a, b = _as_u64(a), _as_u64(b)
        out = np.empty_like(a)
        self.lib.tntt_oracle_schoolbook(self._p(a), self._p(b), self._p(out), a.shape[-1], q)
        return out

    def make_poly(self, seed: int, n: int, q: int) -> np.ndarray:
        out = np.empty(n, dtype=np.uint64)
        self.lib.tntt_oracle_make_poly(seed, n, q, int(q.bit_length() > 32), self._p(out))
        return out

    def checksum(self, v, q: int) -> int:
        v = _as_u64(v)
        return int(self.lib.tntt_oracle_checksum(self._p(v), v.size, int(q.bit_length() > 32)))


def cpu_flags() -> set:
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def best_simd() -> str:
    f = cpu_flags()
    if "avx512f" in f and "avx512dq" in f:
        return "avx512"
    if "avx2" in f:
        return "avx2"
    return "scalar"


class RefLib:
    """The reference's own C++ code for one parameter set (``tag`` in PARAMS)."""

    def __init__(self, tag: str, simd: Optional[str] = None) -> None:
        simd = simd or best_simd()
        path = os.path.join(REF_DIR, f"libref_{tag}_{simd}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.tag, self.simd, self.path = tag, simd, path
        self.lib = L = C.CDLL(path)
        for name in ("tntt_ref_n", "tntt_ref_q", "tntt_ref_psi", "tntt_ref_checksum"):
            getattr(L, name).restype = C.c_uint64
        L.tntt_ref_polymul_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
        L.tntt_ref_forward.argtypes = [C.c_void_p, C.c_void_p]
        L.tntt_ref_schoolbook.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.tntt_ref_make_poly.argtypes = [C.c_uint64, C.c_void_p]
        L.tntt_ref_checksum.argtypes = [C.c_void_p]
        self.n, self.q, self.psi = int(L.tntt_ref_n()), int(L.tntt_ref_q()), int(L.tntt_ref_psi())
        self.dtype = np.uint32 if L.tntt_ref_word_bytes() == 4 else np.uint64
        p = PARAMS[tag]
        assert (self.n, self.q, self.psi) == (p["n"], p["q"], p["psi"]), "reference .so built with other parameters"

    @staticmethod
    def available(tag: str, simd: Optional[str] = None) -> bool:
        return os.path.exists(os.path.join(REF_DIR, f"libref_{tag}_{simd or best_simd()}.so"))

    def _arr(self, x) -> np.ndarray:
        return np.ascontiguousarray(np.asarray(x, dtype=self.dtype))

    def polymul(self, a, b, threads: int = 1) -> np.ndarray:
        a, b = self._arr(a), self._arr(b)
        out = np.empty_like(a)
        self.lib.tntt_ref_polymul_batch(a.ctypes.data, b.ctypes.data, out.ctypes.data, a.size // self.n, threads)
        return out

    def forward(self, a) -> np.ndarray:
        a = self._arr(a)
        out = np.empty_like(a)
        self.lib.tntt_ref_forward(a.ctypes.data, out.ctypes.data)
        return out

    def schoolbook(self, a, b) -> np.ndarray:
        a, b = self._arr(a), self._arr(b)
        out = np.empty_like(a)
        self.lib.tntt_ref_schoolbook(a.ctypes.data, b.ctypes.data, out.ctypes.data)
        return out

    def make_poly(self, seed: int) -> np.ndarray:
        out = np.empty(self.n, dtype=self.dtype)
        self.lib.tntt_ref_make_poly(seed, out.ctypes.data)
        return out

    def checksum(self, v) -> int:
        v = self._arr(v)
        return int(self.lib.tntt_ref_checksum(v.ctypes.data))
