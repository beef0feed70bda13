"""Drop-in for the reference's ``test/refs/ntt_forward_reference.py``:
the parameterised twin of ``cg_ntt`` that the reference's cocotb benches use as their golden model.

Conventions kept (file:line = test/refs/ntt_forward_reference.py):
  * module constants ``N``, ``Q``, ``PSI`` come from the environment variables ``NTT_N``, ``NTT_Q``,
    ``NTT_PSI`` (defaults 4096, 8380417, 283817) and ``OMEGA = PSI^2 mod Q`` (:8-12);
  * ``ntt_forward_reference(coeffs, N=N, q=Q, psi=PSI)`` takes **psi** and transforms with
    ``omega = psi^2`` (:48); inputs are reduced ``% q`` first (:50), natural order in and out;
  * ``ValueError(f"Input must have {N} coefficients, got {n}")`` (:45-46);
  * ``bit_reverse``, ``bit_reverse_list``, ``bit_reverse_order`` (:15-35) are host-side index helpers.
The transform itself runs on the GPU through libtntt.so (``tntt_forward``); there is no CPU path.
"""
import os

from cg_ntt import _download, _plan, _upload, bit_reverse, bit_reverse_list  # noqa: F401
from tntt import ops as _ops

N = int(os.getenv("NTT_N", "4096"))
Q = int(os.getenv("NTT_Q", "8380417"))
PSI = int(os.getenv("NTT_PSI", "283817"))
OMEGA = pow(PSI, 2, Q)


def bit_reverse_order(n):
    """:29-35"""
    bits = (n - 1).bit_length()
    return [bit_reverse(i, bits) for i in range(n)]


def _run(coeffs, N, q, psi, inverse):
    n = len(coeffs)
    if n != N:
        raise ValueError(f"Input must have {N} coefficients, got {n}")
    plan = _plan(N, q, pow(psi, 2, q), False)
    dev = _upload(plan, coeffs, q)
    return _download(plan, _ops.inverse(plan, dev) if inverse else _ops.forward(plan, dev))


def ntt_forward_reference(coeffs, N=N, q=Q, psi=PSI):
    """Constant-geometry cyclic NTT over omega = psi^2, natural order in and out (:38-68)."""
    return _run(coeffs, N, q, psi, False)
