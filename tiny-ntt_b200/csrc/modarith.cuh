// Register-resident modular arithmetic for the sm_100a NTT kernels.
//
// Replaces the reference's arithmetic units:
//   rtl/mod_mult.v:26-135, rtl/barrett_mult.v:36-108, rtl/barrett_reduction.v:23-29  -> barrett_mul()
//   rtl/mod_add.v:14-15, rtl/mod_sub.v:15-17                                         -> lazy add/sub in ct_butterfly()
//   rtl/ntt_butterfly.v:43-72                                                        -> ct_butterfly()
//   software_benchmark/benchmark_ntt.cpp:78-80, benchmark_ntt_60bit.cpp:75-77 (`% Q`) -> shoup_mul()/mont_mul()
//
// Everything is __host__ __device__ so that tests/host_emul can execute the exact
// kernel index maps and arithmetic on the CPU (a test fixture, not a fallback:
// the library proper only ever launches the __global__ kernels).
//
// Lazy-range discipline (Harvey): values live in [0, 2^BITS) between butterflies and
// are only brought back to [0, q) at the final store.  shoup_mul() accepts ANY word
// and returns a value in [0, 2q).  See DESIGN.md "Value ranges".
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define TNTT_HD __host__ __device__ __forceinline__
#else
#define TNTT_HD inline
#endif

namespace tntt {

// ---------------------------------------------------------------- wide multiplies
TNTT_HD uint32_t mulhi(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
TNTT_HD uint64_t mulhi(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

template <typename W> struct WordTraits;
template <> struct WordTraits<uint32_t> { static constexpr int BITS = 32; static constexpr int BANK_BITS = 5; };
template <> struct WordTraits<uint64_t> { static constexpr int BITS = 64; static constexpr int BANK_BITS = 4; };

// twiddle + Shoup companion  wp = floor(w * 2^BITS / q)
template <typename W> struct alignas(2 * sizeof(W)) Tw { W w, wp; };

// ---------------------------------------------------------------- modulus constants
template <typename W> struct Mod {
    W q;        // modulus
    W q2;       // 2q
    W top_sub;  // floor(2^(BITS-1)/q)*q : what csub_top() subtracts
    W nqinv;    // -q^-1 mod 2^BITS  (Montgomery)
    W one_p;    // floor(2^BITS / q) : Shoup companion of 1
    W mu;       // Barrett mu = floor(2^(2k)/q), k = bitlen(q)  (scripts/precompute_constants.py:30-55)
    int k;      // Barrett k
};

// x*w mod q in [0, 2q) for ANY word x; w < q, wp = floor(w*2^BITS/q).   3 / 10 IMAD32
template <typename W> TNTT_HD W shoup_mul(W x, W w, W wp, W q) {
    const W h = mulhi(x, wp);
    return x * w - h * q;
}
template <typename W> TNTT_HD W shoup_mul(W x, const Tw<W> &t, W q) { return shoup_mul(x, t.w, t.wp, q); }

// [0, 2q) -> [0, q)
template <typename W> TNTT_HD W csub(W x, W q) { return x >= q ? x - q : x; }

// any word -> value < max(2^(BITS-1), x - top_sub): keeps lazy values from overflowing.
// Tests only the top bit (one ISETP on the high half for 64-bit words).
TNTT_HD uint32_t csub_top(uint32_t x, uint32_t top_sub) { return ((int32_t)x < 0) ? x - top_sub : x; }
TNTT_HD uint64_t csub_top(uint64_t x, uint64_t top_sub) { return ((int64_t)x < 0) ? x - top_sub : x; }

// Cooley-Tukey butterfly on lazy values: (x, y) <- (x + w*y, x - w*y + 2q).  Grows the bound by 2q.
template <typename W> TNTT_HD void ct_butterfly(W &x, W &y, const Tw<W> &t, const Mod<W> &m) {
    const W v = shoup_mul(y, t.w, t.wp, m.q);
    y = x - v + m.q2;
    x = x + v;
}
// twiddle == 1: v must still be < 2q, so reduce y with the Shoup companion of 1 only when asked
template <typename W> TNTT_HD void ct_butterfly_one(W &x, W &y, const Mod<W> &m) {
    const W v = y - mulhi(y, m.one_p) * m.q;  // y mod q, in [0, 2q)
    y = x - v + m.q2;
    x = x + v;
}

// Montgomery product x*y*2^-BITS mod q, result < x*y/2^BITS + q.
TNTT_HD uint32_t mont_mul(uint32_t x, uint32_t y, const Mod<uint32_t> &m) {
    const uint64_t p = (uint64_t)x * y;
    const uint32_t t = (uint32_t)p * m.nqinv;
    return (uint32_t)((p + (uint64_t)t * m.q) >> 32);
}
TNTT_HD uint64_t mont_mul(uint64_t x, uint64_t y, const Mod<uint64_t> &m) {
    const uint64_t lo = x * y, hi = mulhi(x, y);
    const uint64_t t = lo * m.nqinv;
    return hi + mulhi(t, m.q) + (lo != 0 ? 1u : 0u);  // lo + lo(t*q) == 0 mod 2^64, carry iff lo != 0
}

// The reference's Barrett product for canonical operands (rtl/barrett_reduction.v:23-29):
//   p = a*b; q1 = p >> (k-1); q2 = (q1*mu) >> (k+1); r = p - q2*q; if (r >= q) r -= q
// One conditional subtraction suffices for the shipped moduli; a second
// one is kept so that any q < 2^(BITS-2) is safe.
TNTT_HD uint32_t barrett_mul(uint32_t a, uint32_t b, const Mod<uint32_t> &m) {
    const uint64_t p = (uint64_t)a * b;
    const uint64_t q1 = p >> (m.k - 1);
    const uint64_t q2 = (q1 * m.mu) >> (m.k + 1);   // q1 < 2^(k+1), mu < 2^(k+1): fits for k <= 31
    uint32_t r = (uint32_t)p - (uint32_t)q2 * m.q;
    r = csub(r, m.q);
    return csub(r, m.q);
}
TNTT_HD uint64_t barrett_mul(uint64_t a, uint64_t b, const Mod<uint64_t> &m) {
    const uint64_t lo = a * b, hi = mulhi(a, b);
    const int s1 = m.k - 1;                                   // 1 <= s1 <= 61
    const uint64_t q1 = s1 >= 64 ? 0 : (s1 == 0 ? lo : ((lo >> s1) | (hi << (64 - s1))));  // p < 2^(2k) -> q1 < 2^(k+1)
    const uint64_t l2 = q1 * m.mu, h2 = mulhi(q1, m.mu);
    const int s2 = m.k + 1;                                   // 3 <= s2 <= 63
    const uint64_t q2 = (l2 >> s2) | (h2 << (64 - s2));
    uint64_t r = lo - q2 * m.q;
    r = csub(r, m.q);
    return csub(r, m.q);
}

}  // namespace tntt
