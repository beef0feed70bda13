// Fused negacyclic polymul kernels for uint64 coefficients, N = 4096: the north-star
// configuration q = 2^60 - 2^14 + 1 of rtl/twiddle_forward_4096_60bit.hex (red = 1: lazy
// top-bit reduction before every pass) and any smaller modulus (red = 0).
#include "polymul_inst.cuh"

namespace tntt {
static const PolymulVariant kVariants[] = {
    TNTT_POLYMUL_VARIANT(uint64_t, 64, 12, 4, 1, 1, 1, 2),
    TNTT_POLYMUL_VARIANT(uint64_t, 64, 12, 4, 1, 2, 1, 2),
    TNTT_POLYMUL_VARIANT(uint64_t, 64, 12, 4, 1, 2, 1, 1),
    TNTT_POLYMUL_VARIANT(uint64_t, 64, 12, 3, 1, 1, 1, 2),
    TNTT_POLYMUL_VARIANT(uint64_t, 64, 12, 3, 1, 2, 1, 2),
    TNTT_POLYMUL_VARIANT(uint64_t, 64, 12, 4, 1, 1, 0, 2),
    TNTT_POLYMUL_VARIANT_S(uint64_t, 64, 12, 4, 1, 1, 1, 2, 1),
    TNTT_POLYMUL_VARIANT_S(uint64_t, 64, 12, 4, 1, 1, 1, 3, 1),
    TNTT_POLYMUL_VARIANT(uint64_t, 64, 12, 3, 1, 1, 1, 1),
    TNTT_POLYMUL_VARIANT(uint64_t, 64, 12, 3, 1, 2, 1, 1),
    TNTT_POLYMUL_VARIANT_S(uint64_t, 64, 12, 3, 1, 1, 1, 2, 1),
    // per-thread twiddle tables staged in shared memory by TMA bulk copies
    TNTT_POLYMUL_VARIANT_T(uint64_t, 64, 12, 4, 1, 1, 1, 2, 0, 1),
    TNTT_POLYMUL_VARIANT_T(uint64_t, 64, 12, 4, 1, 1, 1, 1, 1, 1),
    TNTT_POLYMUL_VARIANT_T(uint64_t, 64, 12, 4, 1, 2, 1, 1, 0, 1),
    // round 2: two operands side by side over padded tiles (immediate-offset exchanges), deeper twiddle groups; red2 = the
    // Solinas reductions, matched only by q = 2^60 - 2^14 + 1 (profiles/r02_whatif_*.log)
    TNTT_POLYMUL_VARIANT_P(uint64_t, 64, 12, 4, 1, 2, 1, 2, 0),
    TNTT_POLYMUL_VARIANT_P(uint64_t, 64, 12, 4, 1, 2, 2, 2, 0),
    TNTT_POLYMUL_VARIANT_S(uint64_t, 64, 12, 4, 1, 1, 2, 3, 1),
    TNTT_POLYMUL_VARIANT(uint64_t, 64, 12, 4, 1, 2, 2, 2),
    // red3: the reference's own arithmetic (Barrett products with scripts/precompute_constants.py's k / mu, fully reducing
    // adds) in the same fused kernel -- measured beside the Shoup / Solinas defaults, never a default itself
    TNTT_POLYMUL_VARIANT_S(uint64_t, 64, 12, 4, 1, 1, 3, 3, 1),
    TNTT_POLYMUL_VARIANT_P(uint64_t, 64, 12, 4, 1, 2, 3, 2, 0),
    // small batches: one row per cluster of 4 CTAs, exchanges through distributed shared memory
    TNTT_POLYMUL_CLUSTER(uint64_t, 64, 12, 3, 4, 1),
    TNTT_POLYMUL_CLUSTER(uint64_t, 64, 12, 3, 4, 0),
};
const PolymulVariant *polymul_variants_u64(int *count) {
    *count = (int)(sizeof(kVariants) / sizeof(kVariants[0]));
    return kVariants;
}
}  // namespace tntt
