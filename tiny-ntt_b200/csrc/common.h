// Internal declarations shared by the translation units of libtntt.so.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "kernels.cuh"

namespace tntt {

// one instantiation of the fused polymul kernel
struct PolymulVariant {
    const char *name;
    int word_bytes, logn, logr, ppc, na, red, threads, minb;
    size_t smem;
    // tables / mod point at PolymulTables<W> / Mod<W> of the matching word type
    cudaError_t (*launch)(const void *a, const void *b, void *c, size_t batch, const void *tables, const void *mod,
                          cudaStream_t stream);
    cudaError_t (*prepare)();                                   // opt in to > 48 KB dynamic smem
    cudaError_t (*attributes)(cudaFuncAttributes *attr, int *blocks_per_sm);
    int cluster = 0;   // > 0: one row per thread-block cluster of this many CTAs (polymul_cluster_kernel)
};
// one instantiation of the standalone natural-order transform kernel
struct TransformVariant {
    const char *name;
    int word_bytes, logn, logr, ppc, red, threads;
    size_t smem;
    cudaError_t (*launch)(const void *in, void *out, size_t batch, const void *tables, const void *mod,
                          cudaStream_t stream);
    cudaError_t (*prepare)();
};

// transform-domain kernels of one plan shape (spectrum.cu)
struct SpectrumVariant {
    const char *name;
    int word_bytes, logn, logr, ppc, red;
    cudaError_t (*forward)(const void *in, void *out, size_t batch, const void *tables, const void *mod, cudaStream_t st);
    cudaError_t (*inverse)(const void *in, void *out, size_t batch, const void *tables, const void *post, const void *mod,
                           cudaStream_t st);
    cudaError_t (*polymul)(const void *a, const void *bspec, void *c, size_t batch, size_t b_stride, const void *tables,
                           const void *mod, cudaStream_t st);
    cudaError_t (*prepare)();
    // natural-order transforms built from the same passes (tables decide merged-psi or cyclic)
    cudaError_t (*forward_natural)(const void *in, void *out, size_t batch, const void *tables, const void *mod, cudaStream_t st);
    cudaError_t (*inverse_natural)(const void *in, void *out, size_t batch, const void *tables, const void *post, uint64_t uw,
                                   uint64_t uwp, const void *mod, cudaStream_t st);   // post == nullptr: scale by {uw, uwp}
    int cluster = 0;   // > 0: one row per thread-block cluster of this many CTAs
};
const SpectrumVariant *spectrum_variants(int *count);

const PolymulVariant *polymul_variants_u32(int *count);
const PolymulVariant *polymul_variants_u64(int *count);
const PolymulVariant *polymul_variants_u64b(int *count);
const TransformVariant *transform_variants(int *count);

// element-wise kernels of the literal constant-geometry path (generic.cu); W chosen by word_bytes
cudaError_t launch_cg_stage(int word_bytes, const void *in, void *out, size_t batch, int logn, int stage,
                            const void *pow_table, const void *mod, cudaStream_t stream);
cudaError_t launch_bit_reverse(int word_bytes, const void *in, void *out, size_t batch, int logn, cudaStream_t stream);
cudaError_t launch_pointwise(int word_bytes, const void *a, const void *b, void *c, size_t count, const void *mod,
                             cudaStream_t stream);
cudaError_t launch_mul_table(int word_bytes, const void *in, void *out, size_t batch, int logn, const void *tw_table,
                             const void *mod, cudaStream_t stream);
cudaError_t launch_scale(int word_bytes, const void *in, void *out, size_t count, uint64_t w, uint64_t wp,
                         const void *mod, cudaStream_t stream);
cudaError_t launch_reduce(int word_bytes, const void *in, void *out, size_t count, const void *mod, cudaStream_t stream);

cudaError_t launch_butterfly(const uint64_t *a, const uint64_t *b, const uint64_t *w, uint64_t *out_a, uint64_t *out_b,
                             size_t count, const Mod<uint64_t> &mod, cudaStream_t stream);

cudaError_t run_microbench(int kind, double *ops_per_second);

// records the thread-local message tntt_last_error() returns; returns `code` (capi.cu)
int api_fail(int code, const char *fmt, ...);

}  // namespace tntt
