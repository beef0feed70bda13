"""Drop-in for the reference's ``test/refs/ntt_inverse_reference.py``:
``ntt_inverse_reference(coeffs, N=N, q=Q, psi=PSI)`` = the same constant-geometry schedule with
``omega^-1 = (psi^2)^-1`` followed by the ``N^-1`` scaling (:9-42); same argument checks and error text
as the forward twin.  Runs on the GPU through libtntt.so (``tntt_inverse``).
"""
from .ntt_forward_reference import N, Q, PSI, _run, bit_reverse_list  # noqa: F401


def ntt_inverse_reference(coeffs, N=N, q=Q, psi=PSI):
    return _run(coeffs, N, q, psi, True)
