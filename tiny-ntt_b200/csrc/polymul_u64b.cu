// Fused negacyclic polymul kernels for uint64 coefficients at the smaller sizes (N = 256, 1024).
#include "polymul_inst.cuh"

namespace tntt {
static const PolymulVariant kVariants[] = {
    TNTT_POLYMUL_VARIANT(uint64_t, 64, 8, 4, 16, 1, 0, 2),
    TNTT_POLYMUL_VARIANT(uint64_t, 64, 8, 4, 16, 1, 1, 2),
    TNTT_POLYMUL_VARIANT(uint64_t, 64, 10, 4, 4, 1, 0, 2),
    TNTT_POLYMUL_VARIANT(uint64_t, 64, 10, 4, 4, 1, 1, 2),
};
const PolymulVariant *polymul_variants_u64b(int *count) {
    *count = (int)(sizeof(kVariants) / sizeof(kVariants[0]));
    return kVariants;
}
}  // namespace tntt
