"""Mirror of the reference's ``test/refs`` package surface (test/refs/__init__.py:3-23), GPU-backed."""
from .ntt_forward_reference import (N, OMEGA, PSI, Q, bit_reverse_list, bit_reverse_order,  # noqa: F401
                                    ntt_forward_reference)
from .ntt_inverse_reference import ntt_inverse_reference  # noqa: F401

__all__ = ["N", "Q", "PSI", "OMEGA", "bit_reverse_list", "bit_reverse_order", "ntt_forward_reference",
           "ntt_inverse_reference"]
