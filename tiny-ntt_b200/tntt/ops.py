"""Batched tensor operations over libtntt.so: torch.Tensor / DLPack in, torch.Tensor out.

Rows are polynomials: shape [..., n], CUDA, contiguous, int32|uint32 when plan.word_bytes == 4
and int64|uint64 when it is 8, canonical coefficients (0 <= x < q).  Work is enqueued on the
current torch CUDA stream and is asynchronous like any torch op.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

from . import _lib
from ._lib import TNTT_REDUCE_INPUT, TNTT_TWIST, check, lib
from .plan import Plan


def as_tensor(x):
    """torch.Tensor, or any object exporting DLPack (cupy, jax, numba, ...)."""
    import torch

    if isinstance(x, torch.Tensor):
        return x
    if hasattr(x, "__dlpack__"):
        return torch.from_dlpack(x)
    raise TypeError(f"expected a torch.Tensor or a DLPack exporter, got {type(x).__name__}")


def _prep(plan: Plan, x, name: str):
    import torch

    t = as_tensor(x)
    if not t.is_cuda:
        raise ValueError(f"{name} must live on a CUDA device (no CPU path); use polymul_host for host buffers")
    if t.device.index != plan.device:
        raise ValueError(f"{name} is on cuda:{t.device.index} but the plan is on cuda:{plan.device}")
    if t.dtype not in plan.torch_dtypes:
        raise TypeError(f"{name} has dtype {t.dtype}; this plan (q={plan.q}) uses {plan.torch_dtypes}")
    if t.dim() < 1 or t.shape[-1] != plan.n:
        raise ValueError(f"Expected {plan.n} coefficients, got {t.shape[-1] if t.dim() else 0}")
    if not t.is_contiguous():
        t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone()
    return t


def _out_like(t, out):
    import torch

    if out is None:
        return torch.empty_like(t)
    if out.shape != t.shape or out.dtype != t.dtype or out.device != t.device or not out.is_contiguous():
        raise ValueError("out must match the input's shape, dtype and device and be contiguous")
    return out


def _stream(plan: Plan):
    import torch

    return C.c_void_p(torch.cuda.current_stream(plan.device).cuda_stream)


def _rows(t) -> int:
    return t.numel() // t.shape[-1]


def forward(plan: Plan, x, twist: bool = False, reduce_input: bool = False, out=None):
    """cg_ntt over every row (natural order in and out); twist=True multiplies by psi^i first."""
    t = _prep(plan, x, "x")
    o = _out_like(t, out)
    flags = (TNTT_TWIST if twist else 0) | (TNTT_REDUCE_INPUT if reduce_input else 0)
    check(lib().tntt_forward(plan._h, t.data_ptr(), o.data_ptr(), _rows(t), flags, _stream(plan)))
    return o


def inverse(plan: Plan, x, twist: bool = False, reduce_input: bool = False, out=None):
    """cg_intt over every row (includes N^-1); twist=True also multiplies by psi^-i."""
    t = _prep(plan, x, "x")
    o = _out_like(t, out)
    flags = (TNTT_TWIST if twist else 0) | (TNTT_REDUCE_INPUT if reduce_input else 0)
    check(lib().tntt_inverse(plan._h, t.data_ptr(), o.data_ptr(), _rows(t), flags, _stream(plan)))
    return o


def pointwise(plan: Plan, a, b, out=None):
    ta, tb = _prep(plan, a, "a"), _prep(plan, b, "b")
    if ta.shape != tb.shape:
        raise ValueError("a and b must have the same shape")
    o = _out_like(ta, out)
    check(lib().tntt_pointwise(plan._h, ta.data_ptr(), tb.data_ptr(), o.data_ptr(), _rows(ta), _stream(plan)))
    return o


def polymul(plan: Plan, a, b, out=None, variant: Optional[int] = None):
    """Negacyclic product of every row pair: one fused kernel, one HBM round trip."""
    ta, tb = _prep(plan, a, "a"), _prep(plan, b, "b")
    if ta.shape != tb.shape or ta.dtype != tb.dtype:
        raise ValueError("a and b must have the same shape and dtype")
    o = _out_like(ta, out)
    if variant is None:
        check(lib().tntt_polymul(plan._h, ta.data_ptr(), tb.data_ptr(), o.data_ptr(), _rows(ta), _stream(plan)))
    else:
        check(lib().tntt_polymul_variant(plan._h, variant, ta.data_ptr(), tb.data_ptr(), o.data_ptr(), _rows(ta),
                                         _stream(plan)))
    return o


def forward_spectrum(plan: Plan, x, out=None):
    """Coefficients -> spectrum: ntt(twist(x)) of every row, canonical, in the plan's (opaque) spectrum order.
    Keep operands in this form when they take part in many products (tntt_spectrum_forward)."""
    t = _prep(plan, x, "x")
    o = _out_like(t, out)
    check(lib().tntt_spectrum_forward(plan._h, t.data_ptr(), o.data_ptr(), _rows(t), _stream(plan)))
    return o


def inverse_spectrum(plan: Plan, x, out=None):
    """Spectrum -> coefficients: untwist(cg_intt(.)) of every row (tntt_spectrum_inverse)."""
    t = _prep(plan, x, "x")
    o = _out_like(t, out)
    check(lib().tntt_spectrum_inverse(plan._h, t.data_ptr(), o.data_ptr(), _rows(t), _stream(plan)))
    return o


def polymul_spectrum(plan: Plan, a, b_spectrum, out=None):
    """Negacyclic product a * b with b already in the transform domain (forward_spectrum).  ``b_spectrum`` has
    either a's shape or a single row [n] / [1, n] shared by every row of a.  Bit-identical to polymul(a, b)."""
    ta, tb = _prep(plan, a, "a"), _prep(plan, b_spectrum, "b_spectrum")
    if ta.dtype != tb.dtype:
        raise ValueError("a and b_spectrum must have the same dtype")
    rows, brows = _rows(ta), _rows(tb)
    if brows != rows and brows != 1:
        raise ValueError("b_spectrum must have one row or as many rows as a")
    o = _out_like(ta, out)
    check(lib().tntt_polymul_spectrum(plan._h, ta.data_ptr(), tb.data_ptr(), o.data_ptr(), rows, brows, _stream(plan)))
    return o


def polymul_host(plan: Plan, a, b, out=None):
    """Host tensors in, host tensor out (pin them for full copy/compute overlap).  Blocking."""
    import torch

    ta, tb = as_tensor(a), as_tensor(b)
    if ta.is_cuda or tb.is_cuda:
        raise ValueError("polymul_host takes host tensors")
    if ta.dtype not in plan.torch_dtypes or tb.dtype != ta.dtype:
        raise TypeError(f"host tensors must have dtype in {plan.torch_dtypes}")
    if ta.shape != tb.shape or ta.shape[-1] != plan.n:
        raise ValueError(f"Expected {plan.n} coefficients")
    ta, tb = ta.contiguous(), tb.contiguous()
    o = _host_out(ta, out)
    check(lib().tntt_polymul_host(plan._h, ta.data_ptr(), tb.data_ptr(), o.data_ptr(), _rows(ta)))
    return o


def polymul_spectrum_host(plan: Plan, a, b_spectrum, out=None):
    """polymul_host with the second operand cached on the device: ``a`` is a host tensor (pin it), ``b_spectrum`` a
    CUDA tensor from forward_spectrum with a's number of rows or a single row shared by every row of a; returns a host
    tensor.  Blocking; bit-identical to polymul_host(a, b)."""
    import torch

    ta, tb = as_tensor(a), _prep(plan, b_spectrum, "b_spectrum")
    if ta.is_cuda:
        raise ValueError("polymul_spectrum_host takes a host tensor for a")
    if ta.dtype not in plan.torch_dtypes or tb.dtype != ta.dtype:
        raise TypeError(f"a must be a host tensor with b_spectrum's dtype ({plan.torch_dtypes})")
    if ta.shape[-1] != plan.n:
        raise ValueError(f"Expected {plan.n} coefficients")
    rows, brows = _rows(ta), _rows(tb)
    if brows != rows and brows != 1:
        raise ValueError("b_spectrum must have one row or as many rows as a")
    ta = ta.contiguous()
    o = _host_out(ta, out)
    torch.cuda.current_stream(tb.device).synchronize()   # the spectrum must be complete: the pipeline has its own streams
    check(lib().tntt_polymul_spectrum_host(plan._h, ta.data_ptr(), tb.data_ptr(), brows, o.data_ptr(), rows))
    return o


def _host_out(t, out):
    import torch

    if out is None:
        return torch.empty_like(t, pin_memory=t.is_pinned())
    if out.is_cuda or out.shape != t.shape or out.dtype != t.dtype or not out.is_contiguous():
        raise ValueError("out must be a contiguous host tensor of the input's shape and dtype")
    return out


def polymul_sharded(plans: Sequence[Plan], a, b, out=None):
    """The host-buffer product over several GPUs from this one process (tntt_polymul_host_multi): plans[i] -- plans
    of the same ring on distinct devices, e.g. ``get_plans(n, q, psi)`` -- takes the i-th contiguous range of
    ceil(B / len(plans)) rows; per-device plan, streams and staging buffers, one host thread per device, host
    barrier at the end, no collective.  Host tensors in (pin them), host tensor out."""
    ta, tb = as_tensor(a), as_tensor(b)
    if not plans:
        raise ValueError("need at least one plan")
    plan = plans[0]
    if ta.is_cuda or tb.is_cuda:
        raise ValueError("polymul_sharded takes host tensors")
    if ta.dtype not in plan.torch_dtypes or tb.dtype != ta.dtype:
        raise TypeError(f"host tensors must have dtype in {plan.torch_dtypes}")
    if ta.shape != tb.shape or ta.shape[-1] != plan.n:
        raise ValueError(f"Expected {plan.n} coefficients")
    ta, tb = ta.contiguous(), tb.contiguous()
    o = _host_out(ta, out)
    handles = (C.c_void_p * len(plans))(*[p._h for p in plans])
    check(lib().tntt_polymul_host_multi(handles, len(plans), ta.data_ptr(), tb.data_ptr(), o.data_ptr(), _rows(ta)))
    return o


def cg_stage(plan: Plan, x, stage: int, inverse_root: bool = False, out=None):
    """One literal constant-geometry stage (cg_ntt.py:49-59) -- the verbose/debug path."""
    import torch

    t = _prep(plan, x, "x")
    o = _out_like(t, out)
    if o.data_ptr() == t.data_ptr():
        raise ValueError("cg_stage is out of place: out must not alias x")
    check(lib().tntt_cg_stage(plan._h, t.data_ptr(), o.data_ptr(), _rows(t), stage, int(inverse_root), _stream(plan)))
    return o


def bit_reverse(plan: Plan, x, out=None):
    import torch

    t = _prep(plan, x, "x")
    o = _out_like(t, out)
    if o.data_ptr() == t.data_ptr():
        raise ValueError("bit_reverse is out of place: out must not alias x")
    check(lib().tntt_bit_reverse(plan._h, t.data_ptr(), o.data_ptr(), _rows(t), _stream(plan)))
    return o


def scale(plan: Plan, x, scalar: int, out=None):
    t = _prep(plan, x, "x")
    o = _out_like(t, out)
    check(lib().tntt_scale(plan._h, t.data_ptr(), o.data_ptr(), _rows(t), scalar % plan.q, _stream(plan)))
    return o


def reduce(plan: Plan, x, out=None):
    """x mod q for arbitrary (unsigned) words."""
    t = _prep(plan, x, "x")
    o = _out_like(t, out)
    check(lib().tntt_reduce(plan._h, t.data_ptr(), o.data_ptr(), _rows(t), _stream(plan)))
    return o


def butterfly_lanes(a, b, w, q: int, device: Optional[int] = None):
    """(a + w*b, a - w*b) mod q on every lane; int64 CUDA tensors of canonical values."""
    import torch

    ta, tb, tw = (as_tensor(v).contiguous() for v in (a, b, w))
    for t in (ta, tb, tw):
        if not t.is_cuda or t.dtype not in (torch.int64, torch.uint64) or t.shape != ta.shape:
            raise ValueError("butterfly_lanes needs equally shaped int64 CUDA tensors")
    oa, ob = torch.empty_like(ta), torch.empty_like(ta)
    dev = ta.device.index if device is None else device
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    check(lib().tntt_butterfly_batch(dev, q, ta.data_ptr(), tb.data_ptr(), tw.data_ptr(), oa.data_ptr(), ob.data_ptr(),
                                     ta.numel(), st))
    return oa, ob


def microbench(kind: int, device: int = 0) -> float:
    v = C.c_double()
    check(lib().tntt_microbench(device, kind, C.byref(v)))
    return v.value
