// Element-wise kernels: the literal constant-geometry schedule (one stage per launch), bit
// reversal, pointwise Barrett product, table multiply, scaling, reduction.
//
// These follow the reference's dataflow literally (new_reference/cg_ntt.py:49-59,
// rtl/ntt_cg_address_gen.v:57-117: read (2i, 2i+1), write (i, i+N/2)) with the reference's own
// Barrett reduction (rtl/barrett_reduction.v:23-29).  They serve verbose=True traces and every
// power-of-two n that has no fused kernel; they are HBM-bound streaming kernels.
#include "common.h"

namespace tntt {

static constexpr int kThreads = 256;
static inline unsigned grid_for(size_t work) {
    const size_t g = (work + kThreads - 1) / kThreads;
    return (unsigned)(g < 1 ? 1 : (g > 148u * 64u ? 148u * 64u : g));  // grid-stride beyond 64 CTAs per SM
}

template <typename W>
__global__ void __launch_bounds__(kThreads)
cg_stage_kernel(const W *__restrict__ in, W *__restrict__ out, size_t batch, int logn, int stage,
                const W *__restrict__ pow_table, Mod<W> mod) {
    const size_t half = (size_t)1 << (logn - 1), total = batch * half;
    const int kshift = logn - stage;  // k = n >> stage
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t row = idx >> (logn - 1), i = idx & (half - 1);
        const W *src = in + (row << logn);
        W *dst = out + (row << logn);
        const W w = __ldg(&pow_table[(i >> kshift) << kshift]);  // root^(k * (i / k)), cg_ntt.py:51,54
        const W left = src[2 * i];
        const W t = barrett_mul(w, src[2 * i + 1], mod);
        const W s = left + t;
        dst[i] = s >= mod.q ? s - mod.q : s;                    // rtl/mod_add.v:14-15
        dst[i + half] = left >= t ? left - t : left + mod.q - t;  // rtl/mod_sub.v:15-17
    }
}

template <typename W>
__global__ void __launch_bounds__(kThreads)
bit_reverse_kernel(const W *__restrict__ in, W *__restrict__ out, size_t batch, int logn) {
    const size_t n = (size_t)1 << logn, total = batch * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t row = idx >> logn;
        const unsigned i = (unsigned)(idx & (n - 1));
        const unsigned r = logn ? (__brev(i) >> (32 - logn)) : 0u;
        out[(row << logn) + i] = in[(row << logn) + r];  // gather: coalesced writes
    }
}

template <typename W>
__global__ void __launch_bounds__(kThreads)
pointwise_kernel(const W *a, const W *b, W *c, size_t count, Mod<W> mod) {   // c may alias a or b (in place)
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < count; idx += (size_t)gridDim.x * blockDim.x)
        c[idx] = barrett_mul(a[idx], b[idx], mod);
}

template <typename W>
__global__ void __launch_bounds__(kThreads)
mul_table_kernel(const W *in, W *out, size_t batch, int logn, const Tw<W> *__restrict__ table,   // in == out allowed
                 Mod<W> mod) {
    const size_t n = (size_t)1 << logn, total = batch * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const Tw<W> t = ld_tw(&table[idx & (n - 1)]);
        out[idx] = csub(shoup_mul(in[idx], t.w, t.wp, mod.nq), mod.q);
    }
}

template <typename W>
__global__ void __launch_bounds__(kThreads)
scale_kernel(const W *in, W *out, size_t count, W w, W wp, Mod<W> mod) {   // in == out allowed
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < count; idx += (size_t)gridDim.x * blockDim.x)
        out[idx] = csub(shoup_mul(in[idx], w, wp, mod.nq), mod.q);
}

// 8-lane (or any-lane) butterfly batch of new_reference/cg_ntt_8butterfly.py:8-27 / rtl/ntt_butterfly.v:43-72
__global__ void __launch_bounds__(kThreads)
butterfly_kernel(const uint64_t *a, const uint64_t *b, const uint64_t *w, uint64_t *out_a, uint64_t *out_b, size_t count,
                 Mod<uint64_t> mod) {   // outputs may alias inputs
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < count; idx += (size_t)gridDim.x * blockDim.x) {
        const uint64_t t = barrett_mul(w[idx], b[idx], mod), left = a[idx];
        const uint64_t s = left + t;
        out_a[idx] = s >= mod.q ? s - mod.q : s;
        out_b[idx] = left >= t ? left - t : left + mod.q - t;
    }
}
cudaError_t launch_butterfly(const uint64_t *a, const uint64_t *b, const uint64_t *w, uint64_t *out_a, uint64_t *out_b,
                             size_t count, const Mod<uint64_t> &mod, cudaStream_t st) {
    if (count == 0) return cudaSuccess;
    butterfly_kernel<<<grid_for(count), kThreads, 0, st>>>(a, b, w, out_a, out_b, count, mod);
    return cudaGetLastError();
}

template <typename W> struct Launch {
    static const Mod<W> &M(const void *mod) { return *static_cast<const Mod<W> *>(mod); }
};

#define TNTT_DISPATCH(word_bytes, EXPR32, EXPR64) \
    do { if ((word_bytes) == 4) { EXPR32; } else { EXPR64; } return cudaGetLastError(); } while (0)

cudaError_t launch_cg_stage(int wb, const void *in, void *out, size_t batch, int logn, int stage, const void *pow_table,
                            const void *mod, cudaStream_t st) {
    if (batch == 0) return cudaSuccess;
    const unsigned g = grid_for(batch << (logn - 1));
    TNTT_DISPATCH(wb,
        (cg_stage_kernel<uint32_t><<<g, kThreads, 0, st>>>((const uint32_t *)in, (uint32_t *)out, batch, logn, stage,
                                                           (const uint32_t *)pow_table, Launch<uint32_t>::M(mod))),
        (cg_stage_kernel<uint64_t><<<g, kThreads, 0, st>>>((const uint64_t *)in, (uint64_t *)out, batch, logn, stage,
                                                           (const uint64_t *)pow_table, Launch<uint64_t>::M(mod))));
}
cudaError_t launch_bit_reverse(int wb, const void *in, void *out, size_t batch, int logn, cudaStream_t st) {
    if (batch == 0) return cudaSuccess;
    const unsigned g = grid_for(batch << logn);
    TNTT_DISPATCH(wb,
        (bit_reverse_kernel<uint32_t><<<g, kThreads, 0, st>>>((const uint32_t *)in, (uint32_t *)out, batch, logn)),
        (bit_reverse_kernel<uint64_t><<<g, kThreads, 0, st>>>((const uint64_t *)in, (uint64_t *)out, batch, logn)));
}
cudaError_t launch_pointwise(int wb, const void *a, const void *b, void *c, size_t count, const void *mod, cudaStream_t st) {
    if (count == 0) return cudaSuccess;
    const unsigned g = grid_for(count);
    TNTT_DISPATCH(wb,
        (pointwise_kernel<uint32_t><<<g, kThreads, 0, st>>>((const uint32_t *)a, (const uint32_t *)b, (uint32_t *)c, count,
                                                            Launch<uint32_t>::M(mod))),
        (pointwise_kernel<uint64_t><<<g, kThreads, 0, st>>>((const uint64_t *)a, (const uint64_t *)b, (uint64_t *)c, count,
                                                            Launch<uint64_t>::M(mod))));
}
cudaError_t launch_mul_table(int wb, const void *in, void *out, size_t batch, int logn, const void *table, const void *mod,
                             cudaStream_t st) {
    if (batch == 0) return cudaSuccess;
    const unsigned g = grid_for(batch << logn);
    TNTT_DISPATCH(wb,
        (mul_table_kernel<uint32_t><<<g, kThreads, 0, st>>>((const uint32_t *)in, (uint32_t *)out, batch, logn,
                                                            (const Tw<uint32_t> *)table, Launch<uint32_t>::M(mod))),
        (mul_table_kernel<uint64_t><<<g, kThreads, 0, st>>>((const uint64_t *)in, (uint64_t *)out, batch, logn,
                                                            (const Tw<uint64_t> *)table, Launch<uint64_t>::M(mod))));
}
cudaError_t launch_scale(int wb, const void *in, void *out, size_t count, uint64_t w, uint64_t wp, const void *mod,
                         cudaStream_t st) {
    if (count == 0) return cudaSuccess;
    const unsigned g = grid_for(count);
    TNTT_DISPATCH(wb,
        (scale_kernel<uint32_t><<<g, kThreads, 0, st>>>((const uint32_t *)in, (uint32_t *)out, count, (uint32_t)w,
                                                        (uint32_t)wp, Launch<uint32_t>::M(mod))),
        (scale_kernel<uint64_t><<<g, kThreads, 0, st>>>((const uint64_t *)in, (uint64_t *)out, count, w, wp,
                                                        Launch<uint64_t>::M(mod))));
}
cudaError_t launch_reduce(int wb, const void *in, void *out, size_t count, const void *mod, cudaStream_t st) {
    if (wb == 4) return launch_scale(wb, in, out, count, 1, Launch<uint32_t>::M(mod).one_p, mod, st);
    return launch_scale(wb, in, out, count, 1, Launch<uint64_t>::M(mod).one_p, mod, st);
}

}  // namespace tntt
