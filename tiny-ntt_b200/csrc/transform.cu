// Standalone natural-order transforms (cg_ntt / cg_intt drop-ins, new_reference/cg_ntt.py:29-75):
// instantiations of transform_kernel for the supported sizes.
#include "common.h"

namespace tntt {

template <class C, bool RED, int MINB> struct TransformInst {
    using W = typename C::W;
    static constexpr size_t SMEM = (size_t)C::PPC * C::N * sizeof(W);
    static cudaError_t launch(const void *in, void *out, size_t batch, const void *tables, const void *mod,
                              cudaStream_t stream) {
        if (batch == 0) return cudaSuccess;
        const size_t ctas = (batch + C::PPC - 1) / C::PPC;
        transform_kernel<C, RED, MINB><<<(unsigned)ctas, C::THREADS, SMEM, stream>>>(
            static_cast<const W *>(in), static_cast<W *>(out), batch, *static_cast<const TransformTables<W> *>(tables),
            *static_cast<const Mod<W> *>(mod));
        return cudaGetLastError();
    }
    static cudaError_t prepare() {
        return cudaFuncSetAttribute(transform_kernel<C, RED, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)SMEM);
    }
};

#define TNTT_TRANSFORM_VARIANT(WT, WB, LN, LR, PPC, RED, MINB)                                              \
    TransformVariant {                                                                                      \
        "xf_u" #WB "_n" #LN "_r" #LR "_p" #PPC "_red" #RED, WB / 8, LN, LR, PPC, RED,                          \
            Cfg<WT, LN, LR, PPC>::THREADS, TransformInst<Cfg<WT, LN, LR, PPC>, (RED != 0), MINB>::SMEM,       \
            &TransformInst<Cfg<WT, LN, LR, PPC>, (RED != 0), MINB>::launch,                                  \
            &TransformInst<Cfg<WT, LN, LR, PPC>, (RED != 0), MINB>::prepare                                  \
    }

static const TransformVariant kVariants[] = {
    TNTT_TRANSFORM_VARIANT(uint32_t, 32, 8, 4, 16, 0, 4),
    TNTT_TRANSFORM_VARIANT(uint32_t, 32, 10, 4, 4, 0, 4),
    TNTT_TRANSFORM_VARIANT(uint32_t, 32, 12, 4, 1, 0, 4),
    TNTT_TRANSFORM_VARIANT(uint64_t, 64, 8, 4, 16, 0, 2),
    TNTT_TRANSFORM_VARIANT(uint64_t, 64, 8, 4, 16, 1, 2),
    TNTT_TRANSFORM_VARIANT(uint64_t, 64, 10, 4, 4, 0, 2),
    TNTT_TRANSFORM_VARIANT(uint64_t, 64, 10, 4, 4, 1, 2),
    TNTT_TRANSFORM_VARIANT(uint64_t, 64, 12, 4, 1, 0, 2),
    TNTT_TRANSFORM_VARIANT(uint64_t, 64, 12, 4, 1, 1, 2),
};
const TransformVariant *transform_variants(int *count) {
    *count = (int)(sizeof(kVariants) / sizeof(kVariants[0]));
    return kVariants;
}
}  // namespace tntt
