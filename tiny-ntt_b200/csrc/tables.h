// Host-side constants and twiddle tables of a plan (pure C++, no CUDA).
//
// Replaces the reference's offline generators:
//   scripts/precompute_constants.py:30-55    -> make_mod()  (Barrett k, mu; plus Montgomery / Shoup constants)
//   scripts/generate_twiddles.py:29-41       -> psi_powers(psi)
//   scripts/generate_inverse_twiddles.py:48-61 -> psi_powers(psi^-1)
//   scripts/find_psi.py:9-44                 -> is_primitive_2n_root()
// and arranges the powers in the orders the kernels read them (kernels.cuh).
#pragma once
#include <cstdint>
#include <vector>

#include "modarith.cuh"

namespace tntt {
namespace host {

typedef unsigned __int128 u128;

inline uint64_t mulmod(uint64_t a, uint64_t b, uint64_t q) { return (uint64_t)(((u128)a * b) % q); }
inline uint64_t powmod(uint64_t b, uint64_t e, uint64_t q) {
    uint64_t r = 1 % q;
    b %= q;
    for (; e; e >>= 1, b = mulmod(b, b, q))
        if (e & 1) r = mulmod(r, b, q);
    return r;
}
inline uint64_t modinv(uint64_t v, uint64_t q) { return powmod(v, q - 2, q); }  // new_reference/cg_ntt.py:9-10
inline int ilog2(uint64_t n) { int l = 0; while ((1ull << l) < n) ++l; return l; }
inline int bitlen(uint64_t v) { int l = 0; while (v) { ++l; v >>= 1; } return l; }
inline uint32_t bitrev(uint32_t v, int bits) {
    uint32_t r = 0;
    for (int i = 0; i < bits; ++i) r |= ((v >> i) & 1u) << (bits - 1 - i);
    return r;
}

// deterministic Miller-Rabin for 64-bit integers
inline bool is_prime(uint64_t n) {
    if (n < 2) return false;
    for (uint64_t p : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) {
        if (n % p == 0) return n == p;
    }
    uint64_t d = n - 1; int r = 0;
    while ((d & 1) == 0) { d >>= 1; ++r; }
    for (uint64_t a : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) {
        uint64_t x = powmod(a, d, n);
        if (x == 1 || x == n - 1) continue;
        bool comp = true;
        for (int i = 1; i < r && comp; ++i) { x = mulmod(x, x, n); if (x == n - 1) comp = false; }
        if (comp) return false;
    }
    return true;
}

inline bool is_primitive_2n_root(uint64_t psi, uint64_t n, uint64_t q) {  // scripts/find_psi.py:26-27
    return psi < q && powmod(psi, n, q) == q - 1;
}
inline bool is_primitive_n_root(uint64_t omega, uint64_t n, uint64_t q) {
    return omega < q && (n == 1 ? omega == 1 : powmod(omega, n / 2, q) == q - 1);
}

template <typename W> inline Tw<W> make_tw(uint64_t w, uint64_t q) {
    constexpr int BITS = WordTraits<W>::BITS;
    return Tw<W>{(W)w, (W)((((u128)w) << BITS) / q)};
}

template <typename W> inline bool lazy_full_ok(uint64_t q, int logn);

// `logn` selects the constant of the multiplication-free first inverse stage (Mod::triv_c)
template <typename W> inline Mod<W> make_mod(uint64_t q, int logn = 12) {
    constexpr int BITS = WordTraits<W>::BITS;
    Mod<W> m;
    m.q = (W)q;
    m.nq = (W)(0 - (W)q);
    m.q2 = (W)(2 * q);
    m.qg = (W)(Growth<W>::G * q);
    m.top_sub = (W)((((u128)1 << (BITS - 1)) / q) * q);
    uint64_t inv = q;  // Newton: inv = q^-1 mod 2^64 (q odd)
    for (int i = 0; i < 6; ++i) inv *= 2 - q * inv;
    m.nqinv = (W)(0 - inv);
    m.one_p = (W)(((u128)1 << BITS) / q);
    m.k = bitlen(q);
    m.zero = 0;
    if (lazy_full_ok<W>(q, logn)) {   // inputs of the first inverse stage are below r = fwd^2/2^BITS + q (or 2q)
        const u128 fwd = (u128)q * (1 + Growth<W>::G * logn);
        const u128 r = ((fwd >> 1) * (fwd >> 1) >> (BITS - 2)) + q + 4;
        u128 mult = (r + q - 1) / q;
        if (mult < 2) mult = 2;
        m.triv_c = (W)(mult * q);
    } else {                          // tracked in units of 2^(BITS-4): inputs below 7 units
        const u128 seven = (u128)kTrivMaxIn << (BITS - 4);
        m.triv_c = (W)(((seven + q - 1) / q) * q);
    }
    m.mu = (W)(((u128)1 << (2 * m.k)) / q);
    return m;
}

// table[k] = root^k, k < n   (the rtl/twiddle_*.hex contents)
inline std::vector<uint64_t> powers(uint64_t root, uint32_t n, uint64_t q) {
    std::vector<uint64_t> t(n);
    uint64_t v = 1 % q;
    for (uint32_t k = 0; k < n; ++k) { t[k] = v; v = mulmod(v, root, q); }
    return t;
}

// merged negacyclic Cooley-Tukey twiddles: entry k = psi^bitrev(k, log n)
template <typename W> inline std::vector<Tw<W>> fwd_pyramid(uint64_t psi, uint32_t n, uint64_t q) {
    const int ln = ilog2(n);
    const std::vector<uint64_t> pw = powers(psi, n, q);
    std::vector<Tw<W>> t(n);
    for (uint32_t k = 0; k < n; ++k) t[k] = make_tw<W>(pw[bitrev(k, ln)], q);
    return t;
}

// cyclic Cooley-Tukey twiddles (natural order in, bit-reversed order out): entry m + i, m = 2^s blocks, is
// omega^(bitrev(i, s) * n / 2m) -- the per-block twiddle of cg_ntt's transform when it is run high index bit
// first.  Same indexing as fwd_pyramid, so the same kernels run either transform.
template <typename W> inline std::vector<Tw<W>> fwd_pyramid_cyclic(uint64_t omega, uint32_t n, uint64_t q) {
    const std::vector<uint64_t> pw = powers(omega, n, q);
    std::vector<Tw<W>> t(n > 1 ? n : 2, make_tw<W>(1 % q, q));
    for (uint32_t m = 1, s = 0; m < n; m <<= 1, ++s)
        for (uint32_t i = 0; i < m; ++i) t[m + i] = make_tw<W>(pw[(size_t)bitrev(i, (int)s) * (n / (2 * m))], q);
    return t;
}

// the last forward pass reads fwd_pyramid transposed: [slot][tid], see fwd_pass() in kernels.cuh
template <typename W>
inline std::vector<Tw<W>> fwd_last_table(const std::vector<Tw<W>> &pyr, int logn, int logr) {
    const int R = 1 << logr, P = 1 << (logn - logr);
    std::vector<Tw<W>> t((size_t)(R - 1) * P, Tw<W>{0, 0});
    const int npass = (logn + logr - 1) / logr;
    const int bhi = logn - (npass - 1) * logr;  // the last pass handles index bits [0, bhi)
    for (int b = bhi - 1; b >= 0; --b) {
        const int kb = b, s = logn - 1 - b;
        for (int g = 0; g < (R >> (kb + 1)); ++g)
            for (int tid = 0; tid < P; ++tid)
                t[(size_t)((1 << (logr - 1 - kb)) - 1 + g) * P + tid] = pyr[(1 << s) + (tid << (logr - 1 - kb)) + g];
    }
    return t;
}

// decimation-in-time pyramid of `root`: entry t+j = root^(j * n/(2t)), t = 2^b, j < t
template <typename W> inline std::vector<Tw<W>> dit_pyramid(uint64_t root, uint32_t n, uint64_t q) {
    const int ln = ilog2(n);
    const std::vector<uint64_t> pw = powers(root, n, q);
    std::vector<Tw<W>> t(n > 1 ? n : 2, make_tw<W>(1 % q, q));
    for (int b = 0; b < ln; ++b)
        for (uint32_t j = 0; j < (1u << b); ++j) t[(1u << b) + j] = make_tw<W>(pw[(size_t)j << (ln - 1 - b)], q);
    return t;
}

// entry i = scale * root^i
template <typename W> inline std::vector<Tw<W>> scaled_powers(uint64_t root, uint64_t scale, uint32_t n, uint64_t q) {
    std::vector<Tw<W>> t(n);
    uint64_t v = scale % q;
    for (uint32_t i = 0; i < n; ++i) { t[i] = make_tw<W>(v, q); v = mulmod(v, root, q); }
    return t;
}

// Can all log n stages (plus the pointwise product) run without any intermediate reduction?
//   forward: q(1 + G log n) <= 2^BITS ; Montgomery product r < fwd^2/2^BITS + q ;
//   inverse: r + G q log n <= 2^BITS ; standalone transforms start below 2q: q(2 + G log n) <= 2^BITS
template <typename W> inline bool lazy_full_ok(uint64_t q, int logn) {
    constexpr int BITS = WordTraits<W>::BITS, G = Growth<W>::G;
    const u128 lim = (u128)1 << BITS;
    const u128 fwd = (u128)q * (1 + G * logn);
    if (fwd > lim) return false;
    const u128 r = ((fwd >> 1) * (fwd >> 1) >> (BITS - 2)) + q + 4;  // >= fwd^2 / 2^BITS + q
    // first inverse stage is multiplication-free: (x + y, x - y + c), c = ceil(r/q) q  ->  below 2r + q
    const u128 after_first = 2 * r + q;
    return after_first + (u128)G * (logn - 1) * q <= lim && (u128)q * (4 + G * (logn - 1)) <= lim;
}
// Otherwise the bound tracker of modarith.cuh (units of 2^(BITS-4)) needs q below one unit
template <typename W> inline bool lazy_pass_ok(uint64_t q, int /*logr*/) {
    constexpr int BITS = WordTraits<W>::BITS;
    return q < ((uint64_t)1 << (BITS - 4));
}

}  // namespace host
}  // namespace tntt
