"""Drop-in for the reference's ``new_reference/cg_ntt_8butterfly.py``.

The reference module is the software twin of the RTL's PARALLEL=8 butterfly lanes: identical
results to ``cg_ntt``, issued eight butterflies at a time (cg_ntt_8butterfly.py:41-97).  On the
GPU the lane count is a scheduling detail of the kernels, so the transform entry points are
aliases of the ``cg_ntt`` ones (same bit-exact outputs, same errors, their own log header),
and ``butterfly`` / ``butterfly_batch`` run the butterfly kernel on 1 / 8 lanes.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import cg_ntt as _base
from cg_ntt import N, Q, bit_reverse_list, modinv  # noqa: F401  (same import the reference does, :5)

from tntt import ops as _ops


def _lanes(a_vals, b_vals, omega_vals, modulus: int):
    import numpy as np
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("tiny-ntt_b200 needs a CUDA device; there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device())

    def up(vals):
        arr = np.array([int(v) % modulus for v in vals], dtype=np.uint64).view(np.int64)
        return torch.from_numpy(arr).to(dev)

    try:
        oa, ob = _ops.butterfly_lanes(up(a_vals), up(b_vals), up(omega_vals), modulus, dev.index)
    except _ops._lib.TnttError as exc:
        raise ValueError(exc.message) from None
    to_list = lambda t: [int(v) for v in t.cpu().numpy().view(np.uint64)]  # noqa: E731
    return to_list(oa), to_list(ob)


def butterfly(a: int, b: int, omega: int, modulus: int = Q) -> Tuple[int, int]:
    """(a + omega*b, a - omega*b) mod modulus (:8-10; rtl/ntt_butterfly.v:43-72)."""
    oa, ob = _lanes([a], [b], [omega], modulus)
    return oa[0], ob[0]


def butterfly_batch(a_vals: Sequence[int], b_vals: Sequence[int], omega_vals: Sequence[int],
                    modulus: int = Q) -> Tuple[List[int], List[int]]:
    """Eight butterflies at once (:13-27)."""
    if not (len(a_vals) == len(b_vals) == len(omega_vals) == 8):
        raise ValueError("Expected 8 butterfly lanes")
    return _lanes(a_vals, b_vals, omega_vals, modulus)


def cg_ntt_8butterfly(a_prime, omega_n: int, modulus: int = Q, verbose: bool = False, log_fn=print):
    """:41-97 (this module's own N, imported by value like the reference's, :5)"""
    return _base._transform(N, a_prime, omega_n, modulus, verbose, log_fn, False, "CG NTT 8-butterfly start")


def cg_intt_8butterfly(A, omega_n: int, modulus: int = Q):
    """:100-104"""
    return _base._transform(N, A, omega_n, modulus, False, print, True, "")


def nwc_poly_mult_8butterfly(a, b, psi_2n: int):
    """:107-121"""
    return _base._polymul(N, Q, a, b, psi_2n)
