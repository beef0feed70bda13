"""Plans: ring parameters + device tables (the runtime twin of the reference's compile-time
N/Q/psi and rtl/*.hex tables)."""
from __future__ import annotations

import ctypes as C
import threading
from typing import Dict, List, Optional, Sequence, Tuple

from . import _lib
from ._lib import PlanInfo, TnttError, check, lib


def _require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("tiny-ntt_b200 needs a CUDA device (B200 / sm_100a); there is no CPU fallback")
    return torch


class Plan:
    """One (device, n, q, root) context.  Immutable after creation; safe to share between threads."""

    def __init__(self, handle: int, info: PlanInfo):
        self._h = C.c_void_p(handle)
        self.info = info
        for name, _ in PlanInfo._fields_:
            setattr(self, name, getattr(info, name))

    # -- creation ---------------------------------------------------------------------------
    @classmethod
    def create(cls, n: int, q: int, root: int, root_is_psi: bool = True, device: Optional[int] = None) -> "Plan":
        torch = _require_cuda()
        dev = torch.cuda.current_device() if device is None else int(device)
        h = C.c_void_p()
        check(lib().tntt_plan_create(C.byref(h), dev, n, q, root, 1 if root_is_psi else 0))
        return cls._wrap(h)

    @classmethod
    def from_hex(cls, n: int, q: int, fwd_hex: str, inv_hex: Optional[str] = None, device: Optional[int] = None) -> "Plan":
        """Consume the reference's rtl/twiddle_forward*.hex / twiddle_inverse*.hex tables."""
        torch = _require_cuda()
        dev = torch.cuda.current_device() if device is None else int(device)
        h = C.c_void_p()
        check(lib().tntt_plan_create_from_hex(C.byref(h), dev, n, q, fwd_hex.encode(),
                                              inv_hex.encode() if inv_hex else None))
        return cls._wrap(h)

    @classmethod
    def _wrap(cls, h: C.c_void_p) -> "Plan":
        info = PlanInfo()
        check(lib().tntt_plan_info_get(h, C.byref(info)))
        return cls(h.value, info)

    def write_hex(self, path: str, inverse: bool = False, hex_digits: Optional[int] = None,
                  uppercase: bool = True) -> None:
        """psi^k (or psi^-k) in the $readmemh format of rtl/twiddle_*.hex."""
        digits = hex_digits or (6 if self.q < (1 << 24) else (self.q.bit_length() + 3) // 4)
        check(lib().tntt_plan_write_hex(self._h, path.encode(), int(inverse), digits, int(uppercase)))

    def close(self) -> None:
        if self._h:
            lib().tntt_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # pragma: no cover - interpreter shutdown order
        try:
            self.close()
        except Exception:
            pass

    # -- variants (benchmarking) --------------------------------------------------------------
    def variants(self) -> List[Tuple[int, str]]:
        out = []
        buf = C.create_string_buffer(512)
        for v in range(lib().tntt_variant_count()):
            if lib().tntt_variant_matches(self._h, v):
                check(lib().tntt_variant_describe(v, buf, len(buf)))
                out.append((v, buf.value.decode()))
        return out

    def set_default_variant(self, variant: int) -> None:
        """Benchmarking knob (tntt_plan_set_default_variant): rewrites this plan's dispatch fields and switches the
        batch-size dispatch off.  Not thread-safe, and a plan from get_plan() is shared by every caller of the same ring:
        use it on a private plan (Plan.create) only."""
        check(lib().tntt_plan_set_default_variant(self._h, variant))
        info = PlanInfo()
        check(lib().tntt_plan_info_get(self._h, C.byref(info)))
        self.info = info
        for name, _ in PlanInfo._fields_:
            setattr(self, name, getattr(info, name))

    @property
    def torch_dtypes(self):
        import torch

        return (torch.int32, torch.uint32) if self.word_bytes == 4 else (torch.int64, torch.uint64)

    @property
    def dtype(self):
        """Preferred tensor dtype (signed view: full torch op support; values are < 2^60)."""
        import torch

        return torch.int32 if self.word_bytes == 4 else torch.int64


_cache: Dict[tuple, Plan] = {}
_cache_lock = threading.Lock()


def get_plan(n: int, q: int, root: int, root_is_psi: bool = True, device: Optional[int] = None) -> Plan:
    """Plan cache keyed (device, n, q, root, kind) -- the runtime replacement for the reference's
    module constants / CMake cache variables / Verilog parameters."""
    torch = _require_cuda()
    dev = torch.cuda.current_device() if device is None else int(device)
    key = (dev, int(n), int(q), int(root), bool(root_is_psi))
    with _cache_lock:
        plan = _cache.get(key)
        if plan is None:
            plan = Plan.create(n, q, root, root_is_psi, dev)
            _cache[key] = plan
        return plan


def get_plans(n: int, q: int, root: int, root_is_psi: bool = True, devices: Optional[Sequence[int]] = None) -> List[Plan]:
    """One cached plan per device (default: every visible GPU) for the single-process multi-GPU entry points
    (ops.polymul_sharded / tntt_polymul_host_multi)."""
    torch = _require_cuda()
    devs = list(range(torch.cuda.device_count())) if devices is None else [int(d) for d in devices]
    return [get_plan(n, q, root, root_is_psi, d) for d in devs]


def clear_plan_cache() -> None:
    with _cache_lock:
        for p in _cache.values():
            p.close()
        _cache.clear()
