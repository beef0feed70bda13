"""CPU oracle for the tiny-ntt negacyclic-polymul hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain Python integers, the algorithm of the reference's
golden model so that the CUDA path can be checked bit-for-bit.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it; the product (``tiny-ntt_b200/``) never does and
raises if its CUDA library is missing.

Parity status: PINNED.  ``tests/test_oracle.py`` checks every function below
against (a) the golden vectors in ``tests/golden/`` that were produced by
importing the reference's own ``new_reference/cg_ntt.py`` in the build
container (``tests/golden/make_golden.py``), (b) the reference's known-answer
tests, and (c) the checksums printed by the reference's C++ benchmark binaries
(``software_benchmark/benchmark_ntt{,_60bit}.cpp``) for all four shipped
parameter sets.

Every function cites the reference file:line (relative to /root/reference) it
follows.  Unlike the reference, nothing here depends on module-level N/Q
globals: ring parameters are explicit arguments.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

# ----------------------------------------------------------------------------
# Shipped parameter sets (rtl/*.hex, test/Makefile:264-304,
# software_benchmark/CMakeLists.txt:5-7, rtl/ntt_poly_mult.sv:16-26)
# ----------------------------------------------------------------------------
PARAMS = {
    "dilithium": dict(n=256, q=8380417, psi=1239911),
    "n1024_24": dict(n=1024, q=8380417, psi=5548360),
    "n4096_24": dict(n=4096, q=8380417, psi=283817),
    "n4096_60": dict(n=4096, q=(1 << 60) - (1 << 14) + 1, psi=431606828070683274),
}


# ----------------------------------------------------------------------------
# scalar helpers
# ----------------------------------------------------------------------------
def modinv(value: int, q: int) -> int:
    """Fermat inverse, new_reference/cg_ntt.py:9-10."""
    return pow(value, q - 2, q)


def bit_reverse(value: int, bits: int) -> int:
    """Reverse the low ``bits`` bits of ``value`` (new_reference/cg_ntt.py:13-18)."""
    out = 0
    for pos in range(bits):
        if (value >> pos) & 1:
            out |= 1 << (bits - 1 - pos)
    return out


def bit_reverse_list(values: Sequence[int]) -> List[int]:
    """Scatter values[i] to position bitrev(i) (new_reference/cg_ntt.py:21-26)."""
    bits = (len(values) - 1).bit_length()
    out = [0] * len(values)
    for i, v in enumerate(values):
        out[bit_reverse(i, bits)] = v
    return out


def is_primitive_2n_root(psi: int, n: int, q: int) -> bool:
    """psi^(2n) = 1 and psi^n = -1 (scripts/find_psi.py:26-27, benchmark_ntt.cpp:61-62)."""
    return pow(psi, 2 * n, q) == 1 and pow(psi, n, q) == q - 1


def find_psi(n: int, q: int, max_search: int = 10000) -> Optional[int]:
    """Smallest psi in [2, max_search) that is a primitive 2n-th root (scripts/find_psi.py:9-44)."""
    for cand in range(2, max_search):
        if is_primitive_2n_root(cand, n, q):
            return cand
    return None


# ----------------------------------------------------------------------------
# the constant-geometry transform
# ----------------------------------------------------------------------------
def cg_ntt(
    coeffs: Sequence[int],
    omega: int,
    q: int,
    trace: Optional[Callable[[int, int, int, List[int]], None]] = None,
) -> List[int]:
    """Natural-order cyclic NTT by the Pease constant-geometry schedule.

    Follows new_reference/cg_ntt.py:29-65: bit-reverse the input, then log2(n)
    identical stages reading (2i, 2i+1) and writing (i, i+n/2) with twiddle
    omega^((n>>s) * (i // (n>>s))).  ``trace(stage, k, omega_s, stage_out)`` is
    called after every stage (the reference's verbose log, :60-62).
    Inputs may be any integers; every store is reduced mod q as in :57-59.
    """
    n = len(coeffs)
    half = n // 2
    log_n = (n - 1).bit_length()
    cur = bit_reverse_list(list(coeffs))
    nxt = cur
    for stage in range(1, log_n + 1):
        k = n >> stage
        omega_s = pow(omega, k, q)
        # powers omega_s^j for j = 0 .. (half-1)//k  (reference computes pow() per butterfly, :54)
        pw = [1] * ((half - 1) // k + 1)
        for j in range(1, len(pw)):
            pw[j] = pw[j - 1] * omega_s % q
        nxt = [0] * n
        for i in range(half):
            t = pw[i // k] * cur[2 * i + 1] % q
            nxt[i] = (cur[2 * i] + t) % q
            nxt[i + half] = (cur[2 * i] - t) % q
        if trace is not None:
            trace(stage, k, omega_s, nxt)
        cur = nxt
    return nxt


def cg_intt(values: Sequence[int], omega: int, q: int) -> List[int]:
    """Inverse = forward with omega^-1, then multiply by n^-1 (new_reference/cg_ntt.py:68-75)."""
    n = len(values)
    out = cg_ntt(values, modinv(omega, q), q)
    n_inv = modinv(n, q)
    return [v * n_inv % q for v in out]


def nwc_poly_mult(a: Sequence[int], b: Sequence[int], psi: int, q: int) -> List[int]:
    """Negacyclic product in Z_q[x]/(x^n+1) (new_reference/cg_ntt.py:78-92).

    twist by psi^i, two forward transforms with omega = psi^2, pointwise product,
    inverse transform, untwist by psi^-i.
    """
    n = len(a)
    if len(b) != n:
        raise ValueError("operand lengths differ")
    tw = psi_powers(psi, n, q)
    omega = psi * psi % q
    fa = cg_ntt([x * w % q for x, w in zip(a, tw)], omega, q)
    fb = cg_ntt([x * w % q for x, w in zip(b, tw)], omega, q)
    prod = [x * y % q for x, y in zip(fa, fb)]
    c = cg_intt(prod, omega, q)
    tw_inv = psi_powers(modinv(psi, q), n, q)
    return [x * w % q for x, w in zip(c, tw_inv)]


def forward_negacyclic(a: Sequence[int], psi: int, q: int) -> List[int]:
    """ntt(twist(a)), natural order: what the C++ ``forward_ntt_bench`` produces
    (software_benchmark/benchmark_ntt.cpp:207-211)."""
    n = len(a)
    tw = psi_powers(psi, n, q)
    return cg_ntt([x * w % q for x, w in zip(a, tw)], psi * psi % q, q)


def schoolbook_negacyclic(a: Sequence[int], b: Sequence[int], q: int) -> List[int]:
    """O(n^2) definition of the negacyclic product (new_reference/test_cg_ntt.py:11-21,
    software_benchmark/benchmark_ntt.cpp:213-226)."""
    n = len(a)
    acc = [0] * n
    for i, x in enumerate(a):
        if x == 0:
            continue
        for j, y in enumerate(b):
            d = i + j
            if d < n:
                acc[d] += x * y
            else:
                acc[d - n] -= x * y
    return [v % q for v in acc]


def naive_dft(a: Sequence[int], omega: int, q: int) -> List[int]:
    """X[k] = sum_j a[j] omega^(jk): the function cg_ntt computes."""
    n = len(a)
    return [sum(a[j] * pow(omega, j * k % n, q) for j in range(n)) % q for k in range(n)]


# ----------------------------------------------------------------------------
# tables and constants
# ----------------------------------------------------------------------------
def psi_powers(root: int, n: int, q: int) -> List[int]:
    """table[k] = root^k mod q, k < n (scripts/generate_twiddles.py:29-41,
    scripts/generate_inverse_twiddles.py:48-61 with root = psi^-1)."""
    out = [1] * n
    for k in range(1, n):
        out[k] = out[k - 1] * root % q
    return out


def load_hex_table(path: str) -> List[int]:
    """Parse a $readmemh twiddle file: one upper-case hex word per line
    (scripts/generate_twiddles.py:59-77)."""
    with open(path) as fh:
        return [int(line, 16) for line in fh if line.strip()]


def barrett_constants(q: int) -> Tuple[int, int]:
    """(k, mu) with k = bitlen(q), mu = floor(2^(2k)/q) (scripts/precompute_constants.py:30-55)."""
    k = q.bit_length()
    return k, (1 << (2 * k)) // q


def barrett_reduce(product: int, q: int, k: int, mu: int) -> int:
    """The RTL's reduction (rtl/barrett_reduction.v:23-29): one conditional subtraction."""
    q1 = product >> (k - 1)
    q2 = (q1 * mu) >> (k + 1)
    r = product - q2 * q
    return r - q if r >= q else r


def mod_add(a: int, b: int, q: int) -> int:
    """rtl/mod_add.v:14-15."""
    s = a + b
    return s - q if s >= q else s


def mod_sub(a: int, b: int, q: int) -> int:
    """rtl/mod_sub.v:15-17."""
    return a - b + q if a < b else a - b


def butterfly(a: int, b: int, w: int, q: int) -> Tuple[int, int]:
    """Cooley-Tukey butterfly (a + w*b, a - w*b) (rtl/ntt_butterfly.v:43-72,
    new_reference/cg_ntt_8butterfly.py:8-10)."""
    t = w * b % q
    return (a + t) % q, (a - t) % q


# ----------------------------------------------------------------------------
# the C++ benchmark's input generator and checksum
# ----------------------------------------------------------------------------
_LCG_MUL = 6364136223846793005
_LCG_ADD = 1442695040888963407
_M64 = (1 << 64) - 1
_CK_MOD = 0xFFFFFFFFFFFFFFC5
_CK_MUL = 1315423911


def make_poly24(seed: int, n: int, q: int) -> List[int]:
    """LCG inputs of the 24-bit benchmark: v = (x >> 17) % q (benchmark_ntt.cpp:82-90)."""
    x = seed & _M64
    out = []
    for _ in range(n):
        x = (_LCG_MUL * x + _LCG_ADD) & _M64
        out.append((x >> 17) % q)
    return out


def make_poly60(seed: int, n: int, q: int) -> List[int]:
    """LCG inputs of the 60-bit benchmark: v = x % q (benchmark_ntt_60bit.cpp:79-87)."""
    x = seed & _M64
    out = []
    for _ in range(n):
        x = (_LCG_MUL * x + _LCG_ADD) & _M64
        out.append(x % q)
    return out


def checksum24(values: Sequence[int]) -> int:
    """Fold with uint64 wrap-around BEFORE the modulo (benchmark_ntt.cpp:228-233)."""
    acc = 0
    for v in values:
        acc = ((acc * _CK_MUL + v) & _M64) % _CK_MOD
    return acc


def checksum60(values: Sequence[int]) -> int:
    """Fold widened to 128 bits before the modulo (benchmark_ntt_60bit.cpp:182-188)."""
    acc = 0
    for v in values:
        acc = (acc * _CK_MUL + v) % _CK_MOD
    return acc


def make_poly(tag: str, seed: int) -> List[int]:
    p = PARAMS[tag]
    fn = make_poly60 if p["q"].bit_length() > 32 else make_poly24
    return fn(seed, p["n"], p["q"])


def checksum(tag: str, values: Sequence[int]) -> int:
    return (checksum60 if PARAMS[tag]["q"].bit_length() > 32 else checksum24)(values)
