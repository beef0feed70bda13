// Transform-domain entry points: instantiations of spectrum_forward_kernel / spectrum_inverse_kernel /
// polymul_spectrum_kernel (kernels.cuh) for the plan shapes that have a fused polymul.
#include "common.h"

namespace tntt {

template <class C, int RED, int MINB> struct SpectrumInst {
    using W = typename C::W;
    static constexpr size_t SMEM = (size_t)C::TILE * sizeof(W);
    static unsigned ctas(size_t batch) { return (unsigned)((batch + C::PPC - 1) / C::PPC); }
    static cudaError_t forward(const void *in, void *out, size_t batch, const void *tables, const void *mod, cudaStream_t st) {
        spectrum_forward_kernel<C, RED, MINB><<<ctas(batch), C::THREADS, SMEM, st>>>(
            static_cast<const W *>(in), static_cast<W *>(out), batch, *static_cast<const PolymulTables<W> *>(tables),
            *static_cast<const Mod<W> *>(mod));
        return cudaGetLastError();
    }
    static cudaError_t inverse(const void *in, void *out, size_t batch, const void *tables, const void *post, const void *mod,
                               cudaStream_t st) {
        spectrum_inverse_kernel<C, RED, MINB><<<ctas(batch), C::THREADS, SMEM, st>>>(
            static_cast<const W *>(in), static_cast<W *>(out), batch, static_cast<const PolymulTables<W> *>(tables)->inv,
            static_cast<const Tw<W> *>(post), Tw<W>{0, 0}, *static_cast<const Mod<W> *>(mod));
        return cudaGetLastError();
    }
    static cudaError_t forward_natural(const void *in, void *out, size_t batch, const void *tables, const void *mod, cudaStream_t st) {
        spectrum_forward_kernel<C, RED, MINB, true><<<ctas(batch), C::THREADS, SMEM, st>>>(
            static_cast<const W *>(in), static_cast<W *>(out), batch, *static_cast<const PolymulTables<W> *>(tables),
            *static_cast<const Mod<W> *>(mod));
        return cudaGetLastError();
    }
    static cudaError_t inverse_natural(const void *in, void *out, size_t batch, const void *tables, const void *post, uint64_t uw,
                                       uint64_t uwp, const void *mod, cudaStream_t st) {
        const DitTables<W> &inv = static_cast<const PolymulTables<W> *>(tables)->inv;
        if (post)
            spectrum_inverse_kernel<C, RED, MINB, true, true><<<ctas(batch), C::THREADS, SMEM, st>>>(
                static_cast<const W *>(in), static_cast<W *>(out), batch, inv, static_cast<const Tw<W> *>(post), Tw<W>{0, 0},
                *static_cast<const Mod<W> *>(mod));
        else
            spectrum_inverse_kernel<C, RED, MINB, true, false><<<ctas(batch), C::THREADS, SMEM, st>>>(
                static_cast<const W *>(in), static_cast<W *>(out), batch, inv, nullptr, Tw<W>{(W)uw, (W)uwp},
                *static_cast<const Mod<W> *>(mod));
        return cudaGetLastError();
    }
    static cudaError_t polymul(const void *a, const void *bspec, void *c, size_t batch, size_t b_stride, const void *tables,
                               const void *mod, cudaStream_t st) {
        polymul_spectrum_kernel<C, RED, MINB><<<ctas(batch), C::THREADS, SMEM, st>>>(
            static_cast<const W *>(a), static_cast<const W *>(bspec), static_cast<W *>(c), batch, b_stride,
            *static_cast<const PolymulTables<W> *>(tables), *static_cast<const Mod<W> *>(mod));
        return cudaGetLastError();
    }
    static cudaError_t prepare() {
        cudaError_t e = cudaFuncSetAttribute(spectrum_forward_kernel<C, RED, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(spectrum_inverse_kernel<C, RED, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(polymul_spectrum_kernel<C, RED, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(spectrum_forward_kernel<C, RED, MINB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(spectrum_inverse_kernel<C, RED, MINB, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(spectrum_inverse_kernel<C, RED, MINB, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
        return e;
    }
};

// rows on a thread-block cluster (N = 16384, 32768): spectrum order only, no natural-order kernels
template <class C, int CS, int RED, int MINB> struct SpectrumClusterInst {
    using W = typename C::W;
    static constexpr size_t SMEM = 2 * (size_t)(C::N / CS) * sizeof(W);
    template <int MODE>
    static cudaError_t launch(const void *a, const void *b, void *c, size_t batch, size_t b_stride, const void *tables,
                              const void *post, const void *mod, cudaStream_t st) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(batch * CS));
        cfg.blockDim = dim3(C::P / CS);
        cfg.dynamicSmemBytes = SMEM;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = CS;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, spectrum_cluster_kernel<C, CS, RED, MINB, MODE>, static_cast<const W *>(a),
                                  static_cast<const W *>(b), static_cast<W *>(c), batch, b_stride,
                                  *static_cast<const PolymulTables<W> *>(tables), static_cast<const Tw<W> *>(post),
                                  *static_cast<const Mod<W> *>(mod));
    }
    static cudaError_t forward(const void *in, void *out, size_t batch, const void *tables, const void *mod, cudaStream_t st) {
        return launch<1>(in, nullptr, out, batch, 0, tables, nullptr, mod, st);
    }
    static cudaError_t inverse(const void *in, void *out, size_t batch, const void *tables, const void *post, const void *mod,
                               cudaStream_t st) {
        return launch<2>(in, nullptr, out, batch, 0, tables, post, mod, st);
    }
    static cudaError_t polymul(const void *a, const void *bspec, void *c, size_t batch, size_t b_stride, const void *tables,
                               const void *mod, cudaStream_t st) {
        return launch<3>(a, bspec, c, batch, b_stride, tables, static_cast<const PolymulTables<W> *>(tables)->post, mod, st);
    }
    static cudaError_t prepare() {
        cudaError_t e = cudaFuncSetAttribute(spectrum_cluster_kernel<C, CS, RED, MINB, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(spectrum_cluster_kernel<C, CS, RED, MINB, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(spectrum_cluster_kernel<C, CS, RED, MINB, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
        return e;
    }
};
#define TNTT_SPECTRUM_CLUSTER(WT, WB, LN, LR, CS, RED, MINB)                                                 \
    SpectrumVariant {                                                                                        \
        "sp_u" #WB "_n" #LN "_r" #LR "_red" #RED "_c" #CS, WB / 8, LN, LR, 1, RED,                              \
            &SpectrumClusterInst<Cfg<WT, LN, LR, 1>, CS, RED, MINB>::forward,                           \
            &SpectrumClusterInst<Cfg<WT, LN, LR, 1>, CS, RED, MINB>::inverse,                           \
            &SpectrumClusterInst<Cfg<WT, LN, LR, 1>, CS, RED, MINB>::polymul,                           \
            &SpectrumClusterInst<Cfg<WT, LN, LR, 1>, CS, RED, MINB>::prepare, nullptr, nullptr, CS      \
    }

#define TNTT_SPECTRUM_VARIANT(WT, WB, LN, LR, PPC, RED, MINB) TNTT_SPECTRUM_VARIANT_X("sp_u" #WB "_n" #LN "_r" #LR "_p" #PPC "_red" #RED, WT, WB, LN, LR, PPC, RED, MINB, 0)
// padded tile (kernels.cuh, Cfg): the spectrum order depends on R and P only, so it is the same as without padding
#define TNTT_SPECTRUM_VARIANT_P(WT, WB, LN, LR, PPC, RED, MINB) TNTT_SPECTRUM_VARIANT_X("sp_u" #WB "_n" #LN "_r" #LR "_p" #PPC "_red" #RED "_pad", WT, WB, LN, LR, PPC, RED, MINB, 1)
#define TNTT_SPECTRUM_VARIANT_X(NAME, WT, WB, LN, LR, PPC, RED, MINB, PAD)                                     \
    SpectrumVariant {                                                                                        \
        NAME, WB / 8, LN, LR, PPC, RED,                                                                        \
            &SpectrumInst<Cfg<WT, LN, LR, PPC, PAD>, RED, MINB>::forward,                                      \
            &SpectrumInst<Cfg<WT, LN, LR, PPC, PAD>, RED, MINB>::inverse,                                      \
            &SpectrumInst<Cfg<WT, LN, LR, PPC, PAD>, RED, MINB>::polymul,                                      \
            &SpectrumInst<Cfg<WT, LN, LR, PPC, PAD>, RED, MINB>::prepare,                                      \
            &SpectrumInst<Cfg<WT, LN, LR, PPC, PAD>, RED, MINB>::forward_natural,                              \
            &SpectrumInst<Cfg<WT, LN, LR, PPC, PAD>, RED, MINB>::inverse_natural                               \
    }

// one shape per (word, N, reduction mode): the spectrum order is part of the plan, not of a kernel variant
static const SpectrumVariant kVariants[] = {
    TNTT_SPECTRUM_VARIANT_P(uint32_t, 32, 8, 4, 16, 0, 4),
    TNTT_SPECTRUM_VARIANT_P(uint32_t, 32, 10, 5, 8, 0, 2),
    TNTT_SPECTRUM_VARIANT_P(uint32_t, 32, 12, 4, 1, 0, 4),
    TNTT_SPECTRUM_VARIANT(uint64_t, 64, 8, 4, 16, 0, 2),
    TNTT_SPECTRUM_VARIANT(uint64_t, 64, 8, 4, 16, 1, 2),
    TNTT_SPECTRUM_VARIANT(uint64_t, 64, 10, 4, 4, 0, 2),
    TNTT_SPECTRUM_VARIANT(uint64_t, 64, 10, 4, 4, 1, 2),
    TNTT_SPECTRUM_VARIANT(uint64_t, 64, 12, 4, 1, 0, 2),
    TNTT_SPECTRUM_VARIANT(uint64_t, 64, 12, 4, 1, 2, 3),     // q = 2^60 - 2^14 + 1 only (Solinas reductions): listed before red1
    TNTT_SPECTRUM_VARIANT(uint64_t, 64, 12, 4, 1, 1, 3),
    // N = 512, 2048, 8192
    TNTT_SPECTRUM_VARIANT(uint32_t, 32, 9, 5, 16, 0, 2),
    TNTT_SPECTRUM_VARIANT(uint32_t, 32, 11, 4, 2, 0, 4),
    TNTT_SPECTRUM_VARIANT(uint32_t, 32, 13, 5, 1, 0, 2),
    TNTT_SPECTRUM_VARIANT(uint64_t, 64, 9, 4, 8, 0, 2),
    TNTT_SPECTRUM_VARIANT(uint64_t, 64, 9, 4, 8, 1, 2),
    TNTT_SPECTRUM_VARIANT(uint64_t, 64, 11, 4, 2, 0, 3),
    TNTT_SPECTRUM_VARIANT(uint64_t, 64, 11, 4, 2, 1, 3),
    TNTT_SPECTRUM_VARIANT(uint64_t, 64, 13, 4, 1, 0, 1),
    TNTT_SPECTRUM_VARIANT(uint64_t, 64, 13, 4, 1, 1, 1),
    // N = 16384, 32768: on clusters
    TNTT_SPECTRUM_CLUSTER(uint32_t, 32, 14, 4, 4, 0, 2),
    TNTT_SPECTRUM_CLUSTER(uint32_t, 32, 15, 4, 8, 0, 2),
    TNTT_SPECTRUM_CLUSTER(uint64_t, 64, 14, 4, 4, 0, 2),
    TNTT_SPECTRUM_CLUSTER(uint64_t, 64, 14, 4, 4, 1, 2),
    TNTT_SPECTRUM_CLUSTER(uint64_t, 64, 15, 4, 8, 0, 2),
    TNTT_SPECTRUM_CLUSTER(uint64_t, 64, 15, 4, 8, 1, 2),
};
const SpectrumVariant *spectrum_variants(int *count) {
    *count = (int)(sizeof(kVariants) / sizeof(kVariants[0]));
    return kVariants;
}
}  // namespace tntt
