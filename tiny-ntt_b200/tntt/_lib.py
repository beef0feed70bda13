"""ctypes binding of libtntt.so (include/tntt.h).  There is no CPU fallback: a missing library
or a missing CUDA device is an error."""
from __future__ import annotations

import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# TNTT_LIB_PATH: load another build of the same sources (e.g. the -DTNTT_DEBUG_BOUNDS one, csrc/Makefile `debug`)
LIB_PATH = os.environ.get("TNTT_LIB_PATH") or os.path.join(PKG_DIR, "libtntt.so")

TNTT_OK = 0
TNTT_BAD_ARG, TNTT_BAD_ROOT, TNTT_UNSUPPORTED_N, TNTT_CUDA_ERROR = -1, -2, -3, -4
TNTT_UNSUPPORTED_Q, TNTT_IO_ERROR, TNTT_NO_DEVICE = -5, -6, -7
TNTT_TWIST, TNTT_REDUCE_INPUT = 1, 2

# every symbol include/tntt.h declares (tests check the library exports exactly these)
SYMBOLS = (
    "tntt_plan_create", "tntt_plan_create_from_hex", "tntt_plan_write_hex", "tntt_plan_info_get",
    "tntt_plan_destroy", "tntt_forward", "tntt_inverse", "tntt_pointwise", "tntt_polymul", "tntt_polymul_host",
    "tntt_cg_stage", "tntt_bit_reverse", "tntt_scale", "tntt_reduce", "tntt_butterfly_batch", "tntt_variant_count",
    "tntt_variant_describe", "tntt_variant_matches", "tntt_polymul_variant", "tntt_plan_set_default_variant",
    "tntt_microbench", "tntt_last_error", "tntt_version",
    "tntt_spectrum_forward", "tntt_spectrum_inverse", "tntt_polymul_spectrum", "tntt_plan_info_size",
    "tntt_rns_plan_create", "tntt_rns_plan_destroy", "tntt_rns_plan_limbs", "tntt_rns_plan_word_bytes",
    "tntt_rns_plan_kernel", "tntt_rns_plan_table_bytes", "tntt_rns_polymul", "tntt_rns_plan_check_tables",
    "tntt_rns_kernel_attributes", "tntt_find_psi", "tntt_polymul_host_multi",
    "tntt_rns_spectrum_forward", "tntt_rns_spectrum_inverse", "tntt_rns_polymul_spectrum", "tntt_rns_pointwise",
    "tntt_polymul_spectrum_host",
)


class PlanInfo(C.Structure):
    _fields_ = [
        ("n", C.c_uint32), ("logn", C.c_uint32), ("q", C.c_uint64), ("psi", C.c_uint64), ("psi_inv", C.c_uint64),
        ("omega", C.c_uint64), ("omega_inv", C.c_uint64), ("n_inv", C.c_uint64), ("word_bytes", C.c_int),
        ("barrett_k", C.c_int), ("barrett_mu", C.c_uint64), ("has_psi", C.c_int), ("omega_is_primitive", C.c_int),
        ("fused", C.c_int), ("lazy_reduce", C.c_int), ("default_variant", C.c_int), ("device", C.c_int),
        ("cluster_variant", C.c_int), ("cluster_batch_max", C.c_int), ("small_variant", C.c_int),
        ("small_batch_max", C.c_int), ("spectrum", C.c_int), ("literal_only", C.c_int), ("solinas", C.c_int),
    ]


class TnttError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libtntt error {code}: {message}")
        self.code = code
        self.message = message


_lib = None


def lib() -> C.CDLL:
    """Load libtntt.so (built by tiny-ntt_b200/csrc/Makefile).  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `make -C {os.path.join(PKG_DIR, 'csrc')}` "
            "(there is no CPU fallback for this package)")
    L = C.CDLL(LIB_PATH)
    vp, sz, u64, u32, i = C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint32, C.c_int
    L.tntt_plan_create.argtypes = [C.POINTER(vp), i, u32, u64, u64, i]
    L.tntt_plan_create_from_hex.argtypes = [C.POINTER(vp), i, u32, u64, C.c_char_p, C.c_char_p]
    L.tntt_plan_write_hex.argtypes = [vp, C.c_char_p, i, i, i]
    L.tntt_plan_info_get.argtypes = [vp, C.POINTER(PlanInfo)]
    L.tntt_plan_destroy.argtypes = [vp]
    L.tntt_forward.argtypes = [vp, vp, vp, sz, i, vp]
    L.tntt_inverse.argtypes = [vp, vp, vp, sz, i, vp]
    L.tntt_pointwise.argtypes = [vp, vp, vp, vp, sz, vp]
    L.tntt_polymul.argtypes = [vp, vp, vp, vp, sz, vp]
    L.tntt_polymul_host.argtypes = [vp, vp, vp, vp, sz]
    L.tntt_polymul_host_multi.argtypes = [C.POINTER(vp), i, vp, vp, vp, sz]
    L.tntt_spectrum_forward.argtypes = [vp, vp, vp, sz, vp]
    L.tntt_spectrum_inverse.argtypes = [vp, vp, vp, sz, vp]
    L.tntt_polymul_spectrum.argtypes = [vp, vp, vp, vp, sz, sz, vp]
    L.tntt_polymul_spectrum_host.argtypes = [vp, vp, vp, sz, vp, sz]
    L.tntt_cg_stage.argtypes = [vp, vp, vp, sz, i, i, vp]
    L.tntt_bit_reverse.argtypes = [vp, vp, vp, sz, vp]
    L.tntt_scale.argtypes = [vp, vp, vp, sz, u64, vp]
    L.tntt_reduce.argtypes = [vp, vp, vp, sz, vp]
    L.tntt_butterfly_batch.argtypes = [i, u64, vp, vp, vp, vp, vp, sz, vp]
    L.tntt_variant_count.argtypes = []
    L.tntt_variant_describe.argtypes = [i, C.c_char_p, sz]
    L.tntt_variant_matches.argtypes = [vp, i]
    L.tntt_polymul_variant.argtypes = [vp, i, vp, vp, vp, sz, vp]
    L.tntt_plan_set_default_variant.argtypes = [vp, i]
    L.tntt_microbench.argtypes = [i, i, C.POINTER(C.c_double)]
    L.tntt_rns_plan_create.argtypes = [C.POINTER(vp), i, u32, C.POINTER(u64), C.POINTER(u64), i]
    L.tntt_rns_plan_destroy.argtypes = [vp]
    L.tntt_rns_plan_destroy.restype = None
    L.tntt_rns_plan_limbs.argtypes = L.tntt_rns_plan_word_bytes.argtypes = [vp]
    L.tntt_rns_plan_kernel.argtypes = [vp]
    L.tntt_rns_plan_kernel.restype = C.c_char_p
    L.tntt_rns_plan_table_bytes.argtypes = [vp]
    L.tntt_rns_plan_table_bytes.restype = sz
    L.tntt_rns_polymul.argtypes = [vp, vp, vp, vp, sz, vp]
    L.tntt_rns_plan_check_tables.argtypes = [vp, i]
    L.tntt_rns_spectrum_forward.argtypes = L.tntt_rns_spectrum_inverse.argtypes = [vp, vp, vp, sz, vp]
    L.tntt_rns_polymul_spectrum.argtypes = [vp, vp, vp, vp, sz, sz, vp]
    L.tntt_rns_pointwise.argtypes = [vp, vp, vp, vp, sz, vp]
    L.tntt_rns_kernel_attributes.argtypes = [vp, C.POINTER(i), C.POINTER(sz), C.POINTER(i)]
    L.tntt_find_psi.argtypes = [u32, u64, u64, C.POINTER(u64)]
    L.tntt_last_error.restype = C.c_char_p
    L.tntt_version.restype = i
    L.tntt_plan_info_size.restype = sz
    if L.tntt_plan_info_size() != C.sizeof(PlanInfo):
        raise RuntimeError(f"libtntt.so was built with a different tntt_plan_info ({L.tntt_plan_info_size()} bytes) than "
                           f"this binding mirrors ({C.sizeof(PlanInfo)} bytes): rebuild the library")
    _lib = L
    return L


def check(code: int) -> None:
    if code != TNTT_OK:
        raise TnttError(code, lib().tntt_last_error().decode(errors="replace"))
