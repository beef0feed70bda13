// Instantiation helper for the fused polymul kernel variants.
#pragma once
#include "common.h"

namespace tntt {

template <class C, int NA, int RED, int MINB, int STASH = 0, int TMA = 0> struct PolymulInst {
    using W = typename C::W;
    static constexpr size_t SMEM = ((size_t)NA * C::TILE + (size_t)STASH * C::PPC * C::N) * sizeof(W) + (TMA ? kTwBufBytes + 16 : 0);
    static cudaError_t launch(const void *a, const void *b, void *c, size_t batch, const void *tables, const void *mod,
                              cudaStream_t stream) {
        if (batch == 0) return cudaSuccess;
        const size_t ctas = (batch + C::PPC - 1) / C::PPC;
        polymul_kernel<C, NA, RED, MINB, STASH, TMA><<<(unsigned)ctas, C::THREADS, SMEM, stream>>>(
            static_cast<const W *>(a), static_cast<const W *>(b), static_cast<W *>(c), batch,
            *static_cast<const PolymulTables<W> *>(tables), *static_cast<const Mod<W> *>(mod));
        return cudaGetLastError();
    }
    static cudaError_t prepare() {
        return cudaFuncSetAttribute(polymul_kernel<C, NA, RED, MINB, STASH, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)SMEM);
    }
    static cudaError_t attributes(cudaFuncAttributes *attr, int *blocks_per_sm) {
        cudaError_t e = cudaFuncGetAttributes(attr, polymul_kernel<C, NA, RED, MINB, STASH, TMA>);
        if (e != cudaSuccess) return e;
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, polymul_kernel<C, NA, RED, MINB, STASH, TMA>,
                                                             C::THREADS, SMEM);
    }
};

// one row per cluster of CS CTAs (small batches); launched with a cluster-dimension attribute
template <class C, int CS, int RED, int MINB = 1> struct PolymulClusterInst {
    using W = typename C::W;
    static constexpr size_t SMEM = 2 * (size_t)(C::N / CS) * sizeof(W);
    static cudaError_t launch(const void *a, const void *b, void *c, size_t batch, const void *tables, const void *mod,
                              cudaStream_t stream) {
        if (batch == 0) return cudaSuccess;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(batch * CS));
        cfg.blockDim = dim3(C::P / CS);
        cfg.dynamicSmemBytes = SMEM;
        cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = CS;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, polymul_cluster_kernel<C, CS, RED, MINB>, static_cast<const W *>(a), static_cast<const W *>(b),
                                  static_cast<W *>(c), batch, *static_cast<const PolymulTables<W> *>(tables),
                                  *static_cast<const Mod<W> *>(mod));
    }
    static cudaError_t prepare() {
        return cudaFuncSetAttribute(polymul_cluster_kernel<C, CS, RED, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
    }
    // for cluster variants `blocks_per_sm` reports how many CLUSTERS the device can hold at once (0: the device
    // cannot co-schedule a cluster of this shape, e.g. under a partition with too few SMs per GPC)
    static cudaError_t attributes(cudaFuncAttributes *attr, int *blocks_per_sm) {
        cudaError_t e = cudaFuncGetAttributes(attr, polymul_cluster_kernel<C, CS, RED, MINB>);
        if (e != cudaSuccess) return e;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(CS);
        cfg.blockDim = dim3(C::P / CS);
        cfg.dynamicSmemBytes = SMEM;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = CS;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        return cudaOccupancyMaxActiveClusters(blocks_per_sm, polymul_cluster_kernel<C, CS, RED, MINB>, &cfg);
    }
};
#define TNTT_POLYMUL_CLUSTER(WT, WB, LN, LR, CS, RED) TNTT_POLYMUL_CLUSTER_B(WT, WB, LN, LR, CS, RED, 1)
#define TNTT_POLYMUL_CLUSTER_B(WT, WB, LN, LR, CS, RED, MINB)                                                         \
    PolymulVariant {                                                                                                  \
        "u" #WB "_n" #LN "_r" #LR "_p1_a1_red" #RED "_b" #MINB "_c" #CS, WB / 8, LN, LR, 1, 1, RED,                    \
            Cfg<WT, LN, LR, 1>::P / CS, MINB, PolymulClusterInst<Cfg<WT, LN, LR, 1>, CS, RED, MINB>::SMEM,              \
            &PolymulClusterInst<Cfg<WT, LN, LR, 1>, CS, RED, MINB>::launch,                                            \
            &PolymulClusterInst<Cfg<WT, LN, LR, 1>, CS, RED, MINB>::prepare,                                           \
            &PolymulClusterInst<Cfg<WT, LN, LR, 1>, CS, RED, MINB>::attributes, CS                                     \
    }

// RED: 0 = no intermediate reduction, 1 = lazy top-bit reduction (any q < 2^60), 2 = Solinas forms for q = 2^60 - 2^14 + 1
// PAD: 1 = padded tile instead of the XOR swizzle (64-bit words, 16 coefficients per thread)
#define TNTT_POLYMUL_VARIANT(WT, WB, LN, LR, PPC, NA, RED, MINB) TNTT_POLYMUL_VARIANT_T(WT, WB, LN, LR, PPC, NA, RED, MINB, 0, 0)
#define TNTT_POLYMUL_VARIANT_S(WT, WB, LN, LR, PPC, NA, RED, MINB, ST) TNTT_POLYMUL_VARIANT_T(WT, WB, LN, LR, PPC, NA, RED, MINB, ST, 0)
#define TNTT_POLYMUL_VARIANT_T(WT, WB, LN, LR, PPC, NA, RED, MINB, ST, TM)                                              \
    TNTT_POLYMUL_VARIANT_X("u" #WB "_n" #LN "_r" #LR "_p" #PPC "_a" #NA "_red" #RED "_b" #MINB "_s" #ST "_t" #TM, WT, WB, LN, LR, PPC, \
                           NA, RED, MINB, ST, TM, 0)
#define TNTT_POLYMUL_VARIANT_P(WT, WB, LN, LR, PPC, NA, RED, MINB, ST)                                                  \
    TNTT_POLYMUL_VARIANT_X("u" #WB "_n" #LN "_r" #LR "_p" #PPC "_a" #NA "_red" #RED "_b" #MINB "_s" #ST "_t0_pad", WT, WB, LN, LR, PPC, \
                           NA, RED, MINB, ST, 0, 1)
#define TNTT_POLYMUL_VARIANT_X(NAME, WT, WB, LN, LR, PPC, NA, RED, MINB, ST, TM, PAD)                                   \
    PolymulVariant {                                                                                                  \
        NAME, WB / 8, LN, LR, PPC, NA, RED, Cfg<WT, LN, LR, PPC, PAD>::THREADS, MINB,                                   \
            PolymulInst<Cfg<WT, LN, LR, PPC, PAD>, NA, RED, MINB, ST, TM>::SMEM,                                        \
            &PolymulInst<Cfg<WT, LN, LR, PPC, PAD>, NA, RED, MINB, ST, TM>::launch,                                     \
            &PolymulInst<Cfg<WT, LN, LR, PPC, PAD>, NA, RED, MINB, ST, TM>::prepare,                                    \
            &PolymulInst<Cfg<WT, LN, LR, PPC, PAD>, NA, RED, MINB, ST, TM>::attributes                                  \
    }

}  // namespace tntt
