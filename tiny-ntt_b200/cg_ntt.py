"""Drop-in for the reference's ``new_reference/cg_ntt.py`` -- same names, arguments, return
values and errors -- executing on a B200 through libtntt.so.

    from cg_ntt import N, Q, modinv, bit_reverse, bit_reverse_list, cg_ntt, cg_intt, nwc_poly_mult

Reference conventions kept (all file:line are new_reference/cg_ntt.py):
  * ``N`` and ``Q`` are module globals read at call time, so ``cg_ntt.N = 4096; cg_ntt.Q = q``
    re-targets the module exactly as it does for the reference (:36, :79).  The default
    ``modulus=Q`` argument is bound at import, as in the reference (:32, :68).
  * ``cg_ntt`` / ``cg_intt`` take omega (an N-th root), ``nwc_poly_mult`` takes psi (a 2N-th
    root) and uses the global ``Q`` (:82-92).
  * list in -> new list of Python ints out, natural order, every value in [0, q); inputs are
    not mutated; unreduced or negative integers are accepted (implicit ``%``, :57-59).
  * ``ValueError(f"Expected {N} coefficients, got {len}")`` (:37, :70) and
    ``ValueError(f"Expected {N} coefficients")`` (:80).
  * ``verbose=True`` emits the reference's exact log lines (:43-47, :60-62); that path runs the
    literal constant-geometry schedule one stage per kernel launch (tntt_cg_stage).
Additive: the same functions accept a CUDA ``torch.Tensor`` (or DLPack exporter) of shape
[..., N] and then return a tensor on the same device, computed on the current stream.

There is no CPU fallback: without libtntt.so or a CUDA device these functions raise.
"""
from __future__ import annotations

from typing import List

from tntt import ops as _ops
from tntt._lib import TNTT_BAD_ROOT, TNTT_UNSUPPORTED_N, TNTT_UNSUPPORTED_Q, TnttError as _TnttError
from tntt.plan import get_plan as _get_plan

N = 256
Q = 8380417


def modinv(value: int, modulus: int = Q) -> int:
    """Fermat inverse (:9-10); scalar plan-time constant, evaluated on the host like the reference."""
    return pow(value, modulus - 2, modulus)


def bit_reverse(value: int, bits: int) -> int:
    """:13-18"""
    out = 0
    for _ in range(bits):
        out = (out << 1) | (value & 1)
        value >>= 1
    return out


def bit_reverse_list(values: List[int]) -> List[int]:
    """:21-26 -- a pure index permutation of arbitrary Python objects (host side; the device
    kernels fold the permutation into their load index)."""
    bits = (len(values) - 1).bit_length()
    out = [0] * len(values)
    for idx, val in enumerate(values):
        out[bit_reverse(idx, bits)] = val
    return out


# --------------------------------------------------------------------------------------------
def _is_tensor_like(x) -> bool:
    return not isinstance(x, (list, tuple)) and (hasattr(x, "__dlpack__") or type(x).__name__ == "Tensor")


def _plan(n: int, modulus: int, root: int, is_psi: bool):
    try:
        return _get_plan(n, modulus, root % modulus, is_psi)
    except _TnttError as exc:
        if exc.code in (TNTT_BAD_ROOT, TNTT_UNSUPPORTED_N, TNTT_UNSUPPORTED_Q):
            raise ValueError(exc.message) from None
        raise


def _upload(plan, values, modulus: int):
    """list of Python ints (any size / sign) -> canonical device tensor [1, n]."""
    import numpy as np
    import torch

    np_dtype = np.uint32 if plan.word_bytes == 4 else np.uint64
    arr = np.array([int(v) % modulus for v in values], dtype=np_dtype)
    signed = arr.view(np.int32 if plan.word_bytes == 4 else np.int64)
    return torch.from_numpy(signed).to(torch.device("cuda", plan.device)).unsqueeze(0)


def _download(plan, tensor) -> List[int]:
    import numpy as np

    arr = tensor.reshape(-1).cpu().numpy().view(np.uint32 if plan.word_bytes == 4 else np.uint64)
    return [int(v) for v in arr]


def _transform(n: int, values, omega_n, modulus, verbose, log_fn, inverse: bool, header: str):
    if _is_tensor_like(values):
        x = _ops.as_tensor(values)
        if x.dim() < 1 or x.shape[-1] != n:
            raise ValueError(f"Expected {n} coefficients, got {x.shape[-1] if x.dim() else 0}")
        plan = _plan(n, modulus, omega_n, False)
        return _ops.inverse(plan, x) if inverse else _ops.forward(plan, x)
    if len(values) != n:
        raise ValueError(f"Expected {n} coefficients, got {len(values)}")
    plan = _plan(n, modulus, omega_n, False)
    dev = _upload(plan, values, modulus)
    if inverse:
        return _download(plan, _ops.inverse(plan, dev))
    if not verbose:
        return _download(plan, _ops.forward(plan, dev))
    # verbose: the literal schedule, one stage per launch, logging what the reference logs
    log_fn(header)
    log_fn(f"  omega_n={omega_n} modulus={modulus}")
    log_fn(f"  input(first 16)={list(values[:16])}")
    log_fn(f"  bitrev(first 16)={bit_reverse_list(list(values))[:16]}")
    cur = _ops.bit_reverse(plan, dev)
    log_n = (n - 1).bit_length()
    for stage in range(1, log_n + 1):
        k = n >> stage
        cur = _ops.cg_stage(plan, cur, stage)
        log_fn(f"  stage={stage} k={k} omega_s={pow(omega_n, k, modulus)}")
        log_fn(f"  stage_out(first 16)={_download(plan, cur[..., :16])}")
    return _download(plan, cur)


def cg_ntt(a_prime, omega_n: int, modulus: int = Q, verbose: bool = False, log_fn=print):
    """Natural-order cyclic NTT X[k] = sum_j a[j] omega^(jk) mod q (:29-65)."""
    return _transform(N, a_prime, omega_n, modulus, verbose, log_fn, False, "CG NTT start")


def cg_intt(A, omega_n: int, modulus: int = Q):
    """cg_ntt with omega^-1, then * N^-1 (:68-75)."""
    return _transform(N, A, omega_n, modulus, False, print, True, "")


def nwc_poly_mult(a, b, psi_2n: int):
    """Negacyclic product a*b in Z_Q[x]/(x^N+1) (:78-92): twist, two forward transforms, pointwise
    product, inverse transform and untwist, fused into one kernel with one HBM round trip."""
    return _polymul(N, Q, a, b, psi_2n)


def _polymul(n: int, q: int, a, b, psi_2n: int):
    if _is_tensor_like(a) or _is_tensor_like(b):
        ta, tb = _ops.as_tensor(a), _ops.as_tensor(b)
        if ta.shape[-1] != n or tb.shape[-1] != n:
            raise ValueError(f"Expected {n} coefficients")
        return _ops.polymul(_plan(n, q, psi_2n, True), ta, tb)
    if len(a) != n or len(b) != n:
        raise ValueError(f"Expected {n} coefficients")
    plan = _plan(n, q, psi_2n, True)
    return _download(plan, _ops.polymul(plan, _upload(plan, a, q), _upload(plan, b, q)))
