// Fused negacyclic polymul kernels for uint32 coefficients (q < ~2^26: Dilithium q = 8380417;
// N = 256 / 1024 / 4096 of rtl/twiddle_forward{,_1024,_4096}.hex).
#include "polymul_inst.cuh"

namespace tntt {
static const PolymulVariant kVariants[] = {
    // N = 256: 16 threads x 16 coefficients per polynomial, 16 polynomials per CTA
    TNTT_POLYMUL_VARIANT(uint32_t, 32, 8, 4, 16, 1, 0, 4),
    TNTT_POLYMUL_VARIANT(uint32_t, 32, 8, 4, 16, 2, 0, 4),
    TNTT_POLYMUL_VARIANT(uint32_t, 32, 8, 3, 8, 2, 0, 4),
    // N = 1024: one warp x 32 coefficients, or 64 threads x 16
    TNTT_POLYMUL_VARIANT(uint32_t, 32, 10, 5, 8, 1, 0, 2),
    TNTT_POLYMUL_VARIANT(uint32_t, 32, 10, 5, 8, 2, 0, 2),
    TNTT_POLYMUL_VARIANT(uint32_t, 32, 10, 4, 4, 1, 0, 4),
    TNTT_POLYMUL_VARIANT(uint32_t, 32, 10, 4, 4, 2, 0, 4),
    // N = 4096: one CTA of 256 threads x 16 coefficients
    TNTT_POLYMUL_VARIANT(uint32_t, 32, 12, 4, 1, 1, 0, 4),
    TNTT_POLYMUL_VARIANT(uint32_t, 32, 12, 4, 1, 2, 0, 4),
    TNTT_POLYMUL_VARIANT(uint32_t, 32, 12, 5, 2, 1, 0, 2),
    TNTT_POLYMUL_VARIANT(uint32_t, 32, 12, 3, 1, 2, 0, 2),
    // small batches: one row per cluster of 4 CTAs, exchanges through distributed shared memory
    TNTT_POLYMUL_CLUSTER(uint32_t, 32, 12, 3, 4, 0),
    // three CTAs per SM (80 registers): the four-CTA shapes above spill a little at 64 registers
    TNTT_POLYMUL_VARIANT(uint32_t, 32, 8, 4, 16, 2, 0, 3),
    TNTT_POLYMUL_VARIANT(uint32_t, 32, 12, 4, 1, 2, 0, 3),
    // round 2: padded tiles (immediate-offset exchanges) for the default shapes of the three shipped 24-bit rings
    TNTT_POLYMUL_VARIANT_P(uint32_t, 32, 8, 4, 16, 2, 0, 3, 0),
    TNTT_POLYMUL_VARIANT_P(uint32_t, 32, 10, 5, 8, 2, 0, 2, 0),
    TNTT_POLYMUL_VARIANT_P(uint32_t, 32, 12, 4, 1, 2, 0, 3, 0),
    // sizes next to the reference's three (other NTT-friendly rings): N = 512, 2048, 8192
    TNTT_POLYMUL_VARIANT(uint32_t, 32, 9, 5, 16, 2, 0, 2),
    TNTT_POLYMUL_VARIANT(uint32_t, 32, 11, 4, 2, 2, 0, 3),
    TNTT_POLYMUL_VARIANT(uint32_t, 32, 11, 4, 2, 2, 0, 4),
    TNTT_POLYMUL_VARIANT(uint32_t, 32, 13, 5, 1, 2, 0, 2),
    // rows that do not fit one CTA: one row per thread-block cluster
    TNTT_POLYMUL_CLUSTER_B(uint32_t, 32, 14, 4, 4, 0, 2),
    TNTT_POLYMUL_CLUSTER(uint32_t, 32, 14, 4, 4, 0),
    TNTT_POLYMUL_CLUSTER_B(uint32_t, 32, 15, 4, 8, 0, 2),
    TNTT_POLYMUL_CLUSTER(uint32_t, 32, 15, 4, 8, 0),
};
const PolymulVariant *polymul_variants_u32(int *count) {
    *count = (int)(sizeof(kVariants) / sizeof(kVariants[0]));
    return kVariants;
}
}  // namespace tntt
