"""Loader for the host emulation of the kernels (tests/host_emul.cpp).  Test fixture only."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libhost_emul.so")
SRC = os.path.join(HERE, "host_emul.cpp")
CSRC = os.path.join(os.path.dirname(HERE), "tiny-ntt_b200", "csrc")

_lib = None


def lib():
    global _lib
    if _lib is None:
        deps = [SRC] + [os.path.join(CSRC, f) for f in ("kernels.cuh", "modarith.cuh", "tables.h")]
        if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(d) for d in deps):
            os.makedirs(os.path.dirname(SO), exist_ok=True)
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", SO, SRC])
        L = C.CDLL(SO)
        L.emu_polymul.argtypes = [C.c_int] * 6 + [C.c_void_p] * 3 + [C.c_size_t, C.c_uint64, C.c_uint64]
        L.emu_transform.argtypes = [C.c_int] * 5 + [C.c_void_p] * 2 + [C.c_size_t, C.c_uint64, C.c_uint64, C.c_int, C.c_int]
        L.emu_slot.argtypes = [C.c_int] * 7
        L.emu_polymul_ex.argtypes = [C.c_int] * 7 + [C.c_void_p] * 3 + [C.c_size_t, C.c_uint64, C.c_uint64]
        L.emu_slot_ex.argtypes = [C.c_int] * 8
        L.emu_cluster_exchange_violations.argtypes = [C.c_int] * 5
        L.emu_solinas_reduce.argtypes = [C.c_uint64]
        L.emu_solinas_reduce.restype = C.c_uint64
        L.emu_solinas_mul.argtypes = [C.c_uint64] * 2
        L.emu_solinas_mul.restype = C.c_uint64
        L.emu_spectrum.argtypes = [C.c_int] * 5 + [C.c_void_p] * 3 + [C.c_size_t, C.c_uint64, C.c_uint64, C.c_int]
        for name, t in (("emu_shoup64", C.c_uint64), ("emu_shoup_lazy64", C.c_uint64), ("emu_mont64", C.c_uint64), ("emu_barrett64", C.c_uint64),
                        ("emu_csub_top64", C.c_uint64), ("emu_shoup32", C.c_uint32), ("emu_mont32", C.c_uint32),
                        ("emu_barrett32", C.c_uint32)):
            getattr(L, name).restype = t
        L.emu_shoup64.argtypes = L.emu_shoup_lazy64.argtypes = L.emu_mont64.argtypes = L.emu_barrett64.argtypes = [C.c_uint64] * 3
        L.emu_shoup32.argtypes = L.emu_mont32.argtypes = L.emu_barrett32.argtypes = [C.c_uint32] * 3
        L.emu_csub_top64.argtypes = [C.c_uint64] * 2
        L.emu_is_prime.argtypes = [C.c_uint64]
        L.emu_range_violations.restype = C.c_longlong
        L.emu_lazy_full_ok.argtypes = [C.c_int, C.c_uint64, C.c_int]
        L.emu_dit2_pass0_bounds.argtypes = [C.c_int] * 4 + [C.POINTER(C.c_int)]
        L.emu_dit2_bound_at.argtypes = [C.c_int] * 4
        L.emu_dit2_step.argtypes = [C.c_int] * 4 + [C.POINTER(C.c_int)]
        _lib = L
    return _lib


def polymul(wb, logn, logr, ppc, na, red, a, b, q, psi, pad=0):
    """red: 0 / 1 / 2 (Solinas reductions, q = 2^60 - 2^14 + 1 only); pad = 1: padded tile instead of the XOR swizzle"""
    dt = np.uint32 if wb == 4 else np.uint64
    a = np.ascontiguousarray(a, dtype=dt)
    b = np.ascontiguousarray(b, dtype=dt)
    c = np.zeros_like(a)
    rc = lib().emu_polymul_ex(wb, logn, logr, ppc, na, red, pad, a.ctypes.data, b.ctypes.data, c.ctypes.data, a.size >> logn, q, psi)
    if rc:
        raise RuntimeError(f"emu_polymul rc={rc}")
    return c


def spectrum(wb, logn, logr, ppc, red, a, b, q, psi, mode):
    """mode 0: inverse(forward(a)); 1: polymul_spectrum(a, forward(b)); 2: same, b[0]'s spectrum shared; 3: forward(a)"""
    dt = np.uint32 if wb == 4 else np.uint64
    a = np.ascontiguousarray(a, dtype=dt)
    b = np.ascontiguousarray(b, dtype=dt)
    out = np.zeros_like(a)
    rc = lib().emu_spectrum(wb, logn, logr, ppc, red, a.ctypes.data, b.ctypes.data, out.ctypes.data, a.size >> logn, q, psi, mode)
    if rc:
        raise RuntimeError(f"emu_spectrum rc={rc}")
    return out


def transform(wb, logn, logr, ppc, red, x, q, root, mode, reduce_input=0):
    dt = np.uint32 if wb == 4 else np.uint64
    x = np.ascontiguousarray(x, dtype=dt)
    out = np.zeros_like(x)
    rc = lib().emu_transform(wb, logn, logr, ppc, red, x.ctypes.data, out.ctypes.data, x.size >> logn, q, root, mode,
                             reduce_input)
    if rc:
        raise RuntimeError(f"emu_transform rc={rc}")
    return out


def largest_friendly_prime(n, ok):
    """Largest prime q = 1 (mod 2n) below 2^60 for which ok(q) holds (ok is monotone: true below a limit)."""
    L = lib()
    lo, hi = 3, (1 << 60) - 1
    while lo < hi:                                   # largest q with ok(q)
        mid = (lo + hi + 1) // 2
        lo, hi = (mid, hi) if ok(mid) else (lo, mid - 1)
    q = lo - (lo - 1) % (2 * n)                      # largest value = 1 mod 2n not above the limit
    while not L.emu_is_prime(q):
        q -= 2 * n
    return q


def largest_modulus_of_path(word_bytes, logn, red):
    """The largest NTT-friendly prime the (word, reduction mode) kernel path of size 2^logn admits."""
    L = lib()
    ok = (lambda q: q < (1 << 60)) if red else (lambda q: bool(L.emu_lazy_full_ok(word_bytes, q, logn)))
    return largest_friendly_prime(1 << logn, ok)
