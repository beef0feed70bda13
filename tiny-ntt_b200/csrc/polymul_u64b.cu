// Fused negacyclic polymul kernels for uint64 coefficients at the sizes other than N = 4096.
#include "polymul_inst.cuh"

namespace tntt {
static const PolymulVariant kVariants[] = {
    TNTT_POLYMUL_VARIANT(uint64_t, 64, 8, 4, 16, 1, 0, 2),
    TNTT_POLYMUL_VARIANT(uint64_t, 64, 8, 4, 16, 1, 1, 2),
    TNTT_POLYMUL_VARIANT(uint64_t, 64, 10, 4, 4, 1, 0, 2),
    TNTT_POLYMUL_VARIANT(uint64_t, 64, 10, 4, 4, 1, 1, 2),
    // sizes next to the reference's three (other NTT-friendly rings): N = 512, 2048, 8192
    TNTT_POLYMUL_VARIANT(uint64_t, 64, 9, 4, 8, 1, 0, 2),
    TNTT_POLYMUL_VARIANT(uint64_t, 64, 9, 4, 8, 1, 1, 2),
    TNTT_POLYMUL_VARIANT_S(uint64_t, 64, 11, 4, 2, 1, 0, 3, 1),
    TNTT_POLYMUL_VARIANT(uint64_t, 64, 11, 4, 2, 2, 1, 2),          // first match = default (25.2 vs 24.1 M/s measured)
    TNTT_POLYMUL_VARIANT_S(uint64_t, 64, 11, 4, 2, 1, 1, 3, 1),
    TNTT_POLYMUL_VARIANT_S(uint64_t, 64, 13, 4, 1, 1, 0, 1, 1),
    TNTT_POLYMUL_VARIANT_S(uint64_t, 64, 13, 4, 1, 1, 1, 1, 1),
    // rows that do not fit one CTA: one row per thread-block cluster (4 / 8 CTAs of 256 threads x 16 coefficients),
    // every exchange through distributed shared memory.  Two CTAs (of different clusters) per SM hide each other's
    // cluster barriers: 1.51 vs 1.24 M polymul/s at N = 16384 / 60-bit although the 128-register shape spills.
    TNTT_POLYMUL_CLUSTER_B(uint64_t, 64, 14, 4, 4, 0, 2),
    TNTT_POLYMUL_CLUSTER_B(uint64_t, 64, 14, 4, 4, 1, 2),
    TNTT_POLYMUL_CLUSTER(uint64_t, 64, 14, 4, 4, 1),
    TNTT_POLYMUL_CLUSTER_B(uint64_t, 64, 15, 4, 8, 0, 2),
    TNTT_POLYMUL_CLUSTER_B(uint64_t, 64, 15, 4, 8, 1, 2),
    TNTT_POLYMUL_CLUSTER(uint64_t, 64, 15, 4, 8, 1),
};
const PolymulVariant *polymul_variants_u64b(int *count) {
    *count = (int)(sizeof(kVariants) / sizeof(kVariants[0]));
    return kVariants;
}
}  // namespace tntt
