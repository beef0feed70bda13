"""tntt -- host side of the B200 batched negacyclic-polymul engine (libtntt.so over ctypes).

The reference-compatible entry points live one level up, in ``cg_ntt.py`` and
``cg_ntt_8butterfly.py`` (same names and conventions as the reference's
new_reference/ modules); this package holds the binding, the plan cache and the
batched tensor API they are built on.
"""
from . import fixtures  # noqa: F401
from ._lib import LIB_PATH, SYMBOLS, TnttError, lib  # noqa: F401
from .ops import (as_tensor, bit_reverse, butterfly_lanes, cg_stage, forward, forward_spectrum, inverse,  # noqa: F401
                  inverse_spectrum, microbench, pointwise, polymul, polymul_host, polymul_sharded, polymul_spectrum, polymul_spectrum_host,
                  reduce,
                  scale)
from .plan import Plan, clear_plan_cache, get_plan, get_plans  # noqa: F401
from .rns import RnsContext, find_psi  # noqa: F401
from .shard import shard_range, shard_rows  # noqa: F401

__all__ = [
    "Plan", "get_plan", "clear_plan_cache", "forward", "inverse", "pointwise", "polymul", "polymul_host", "cg_stage",
    "bit_reverse", "scale", "reduce", "butterfly_lanes", "microbench", "shard_range", "shard_rows", "TnttError", "lib",
    "RnsContext", "find_psi", "fixtures", "forward_spectrum", "inverse_spectrum", "polymul_spectrum", "polymul_sharded",
    "get_plans", "polymul_spectrum_host",
]
