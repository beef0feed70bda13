#!/usr/bin/env python3
"""Does write-combined pinned memory raise host-to-device bandwidth when several GPUs pull at once?
Run under torchrun (one rank per GPU); every rank copies 256 MiB chunks H2D from normal pinned and from
write-combined pinned memory (cudaHostAlloc through libcudart), alone and concurrently."""
import ctypes as C
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("gloo")
rt = C.CDLL("libcudart.so.12")
n = 256 << 20
dev = torch.empty(n, dtype=torch.uint8, device="cuda")


def host_alloc(flags):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(n), C.c_uint(flags)) == 0
    C.memset(p, 1, n)
    return p


def bw(ptr, reps=8):
    rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    rt.cudaMemcpyAsync(C.c_void_p(dev.data_ptr()), ptr, n, 1, st); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t = time.perf_counter()
    for _ in range(reps): rt.cudaMemcpyAsync(C.c_void_p(dev.data_ptr()), ptr, n, 1, st)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    if world > 1: dist.barrier()
    return n * reps / dt / 1e9


def bw_d2h(ptr, reps=8, also_h2d=None):
    rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t = time.perf_counter()
    for _ in range(reps):
        rt.cudaMemcpyAsync(ptr, C.c_void_p(dev.data_ptr()), n, 2, C.c_void_p(s1.cuda_stream))
        if also_h2d is not None:
            rt.cudaMemcpyAsync(C.c_void_p(dev2.data_ptr()), also_h2d, n, 1, C.c_void_p(s2.cuda_stream))
            rt.cudaMemcpyAsync(C.c_void_p(dev2.data_ptr()), also_h2d, n, 1, C.c_void_p(s2.cuda_stream))
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    if world > 1: dist.barrier()
    return n * reps / dt / 1e9


dev2 = torch.empty(n, dtype=torch.uint8, device="cuda")
normal, wc = host_alloc(0), host_alloc(4)   # cudaHostAllocDefault, cudaHostAllocWriteCombined
out_n, out_wc = host_alloc(0), host_alloc(4)
for name, ptr in (("pinned", normal), ("write-combined", wc), ("pinned", normal), ("write-combined", wc)):
    v = bw(ptr)
    print(f"rank {rank}/{world} H2D from {name}: {v:.1f} GB/s", flush=True)
for name, ptr in (("pinned", out_n), ("write-combined", out_wc)):
    print(f"rank {rank}/{world} D2H to {name}: {bw_d2h(ptr):.1f} GB/s", flush=True)
for name, ptr, src in (("pinned", out_n, normal), ("write-combined", out_wc, wc), ("pinned out / wc in", out_n, wc)):
    v = bw_d2h(ptr, also_h2d=src)
    print(f"rank {rank}/{world} 1 D2H + 2 H2D ({name}): D2H {v:.1f} GB/s, H2D {2 * v:.1f} GB/s", flush=True)
