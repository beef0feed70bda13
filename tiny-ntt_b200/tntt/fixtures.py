"""Input generator and checksum of the reference's C++ benchmark.

``make_poly`` / ``checksum`` are what `software_benchmark/benchmark_ntt.cpp:82-90,228-233` (24-bit
build) and `software_benchmark/benchmark_ntt_60bit.cpp:79-87,182-188` (60-bit build) use to print
their ``checksum=`` line; a caller that switches from the reference binary to this engine needs the
same two functions to reproduce that line.  They are fixture helpers: plain integer loops on the
host, no device work, and they are not an alternative compute path (there is none).

The two builds differ: the 24-bit one draws ``(x >> 17) % q`` and folds its checksum with uint64
wrap-around *before* the modulo; the 60-bit one draws ``x % q`` and widens the fold to 128 bits.
"""
from __future__ import annotations

from typing import List, Sequence

_LCG_MUL = 6364136223846793005
_LCG_ADD = 1442695040888963407
_M64 = (1 << 64) - 1
_CK_MOD = 0xFFFFFFFFFFFFFFC5
_CK_MUL = 1315423911

# `checksum=` lines printed by the reference binaries for make_poly(1) * make_poly(2)
REFERENCE_CHECKSUMS = {
    (256, 8380417): 16424788039373839479,
    (1024, 8380417): 15308795525113097448,
    (4096, 8380417): 11303505593119465445,
    (4096, (1 << 60) - (1 << 14) + 1): 2710933653778106521,
}


def make_poly(seed: int, n: int, q: int) -> List[int]:
    """LCG polynomial of the benchmark that is built for modulus ``q`` (60-bit build when q >= 2^32)."""
    wide = q.bit_length() > 32
    x = seed & _M64
    out = []
    for _ in range(n):
        x = (_LCG_MUL * x + _LCG_ADD) & _M64
        out.append(x % q if wide else (x >> 17) % q)
    return out


def checksum(values: Sequence[int], q: int) -> int:
    """The benchmark's fold of a result polynomial (the value after ``checksum=``)."""
    wide = q.bit_length() > 32
    acc = 0
    for v in values:
        acc = (acc * _CK_MUL + int(v)) % _CK_MOD if wide else ((acc * _CK_MUL + int(v)) & _M64) % _CK_MOD
    return acc
