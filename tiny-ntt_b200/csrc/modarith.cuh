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
#define TNTT_CX __host__ __device__ constexpr
#else
#define TNTT_HD inline
#define TNTT_CX constexpr
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
    W nq;       // 2^BITS - q  (so that x*w - h*q == x*w + h*nq without a negation)
    W q2;       // 2q
    W qg;       // GROWTH*q: what a butterfly adds to keep x - w*y non-negative (2q for 32-bit, 3q for 64-bit words)
    W top_sub;  // floor(2^(BITS-1)/q)*q : what csub_top() subtracts
    W nqinv;    // -q^-1 mod 2^BITS  (Montgomery)
    W one_p;    // floor(2^BITS / q) : Shoup companion of 1
    W mu;       // Barrett mu = floor(2^(2k)/q), k = bitlen(q)  (scripts/precompute_constants.py:30-55)
    W triv_c;   // multiple of q added by the multiplication-free first DIT stage: >= the bound of that stage's inputs
    W zero;     // always 0.  Added as a third operand to two-input 64-bit additions: a three-input add can only
                // issue as IADD3 on the ALU pipe, which stops ptxas from turning it into IMAD.X on the
                // (already saturated) multiplier pipe.  It travels in the constant bank, so it costs nothing.
    int k;      // Barrett k
};

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void unpack64(uint64_t v, uint32_t &lo, uint32_t &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t pack64(uint32_t lo, uint32_t hi) {
    uint64_t v;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "r"(lo), "r"(hi));
    return v;
}
#endif

// x*w mod q in [0, 2q) for ANY word x; w < q, wp = floor(w*2^BITS/q), nq = 2^BITS - q.
// 32-bit: 3 IMAD.  64-bit: 10 IMAD + 3 carry ops = 13 SASS instructions -- the low 64 bits of
// x*w + h*nq are accumulated through one IMAD.WIDE chain and four IMAD.LO into the high word
// (pinned with PTX: left to itself nvcc builds both 64-bit products separately and negates).
TNTT_HD uint32_t shoup_mul(uint32_t x, uint32_t w, uint32_t wp, uint32_t nq) {
    const uint32_t h = mulhi(x, wp);
    return x * w + h * nq;
}
TNTT_HD uint64_t shoup_mul(uint64_t x, uint64_t w, uint64_t wp, uint64_t nq) {
    const uint64_t h = mulhi(x, wp);
#if defined(__CUDA_ARCH__)
    uint32_t x0, x1, w0, w1, h0, h1, n0, n1, lo, hi;
    unpack64(x, x0, x1);
    unpack64(w, w0, w1);
    unpack64(h, h0, h1);
    unpack64(nq, n0, n1);
    uint64_t acc;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(acc) : "r"(x0), "r"(w0));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(h0), "r"(n0));
    unpack64(acc, lo, hi);
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(x0), "r"(w1));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(x1), "r"(w0));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(h0), "r"(n1));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(h1), "r"(n0));
    return pack64(lo, hi);
#else
    return x * w + h * nq;
#endif
}
template <typename W> TNTT_HD W shoup_mul(W x, const Tw<W> &t, W nq) { return shoup_mul(x, t.w, t.wp, nq); }

// [0, 2q) -> [0, q)
template <typename W> TNTT_HD W csub(W x, W q) { return x >= q ? x - q : x; }

// any word -> value < max(2^(BITS-1), x - top_sub): keeps lazy values from overflowing.
// Tests only the top bit (one ISETP on the high half for 64-bit words).
TNTT_HD uint32_t csub_top(uint32_t x, uint32_t top_sub) { return ((int32_t)x < 0) ? x - top_sub : x; }
TNTT_HD uint64_t csub_top(uint64_t x, uint64_t top_sub) {
#if defined(__CUDA_ARCH__)
    uint32_t lo, hi, c0, c1;
    unpack64(x, lo, hi);
    unpack64(top_sub, c0, c1);
    asm("{\n\t.reg .pred p;\n\tsetp.lt.s32 p, %1, 0;\n\t@p sub.cc.u32 %0, %0, %2;\n\t@p subc.u32 %1, %1, %3;\n\t}"
        : "+r"(lo), "+r"(hi) : "r"(c0), "r"(c1));
    return pack64(lo, hi);
#else
    return ((int64_t)x < 0) ? x - top_sub : x;
#endif
}

// Host-only audit used by tests/host_emul.cpp: counts lazy values that wrapped around 2^BITS.
#if !defined(__CUDA_ARCH__) && defined(TNTT_AUDIT_RANGES)
extern long long g_range_violations;
template <typename W> inline void audit_butterfly(W x, W v, W qg) {
    const unsigned __int128 lim = (unsigned __int128)1 << WordTraits<W>::BITS;
    if ((unsigned __int128)x + v >= lim || (unsigned __int128)x + qg >= lim + v || v >= qg) ++g_range_violations;
}
#define TNTT_AUDIT_BUTTERFLY(x, v, qg) audit_butterfly(x, v, qg)
#else
#define TNTT_AUDIT_BUTTERFLY(x, v, qg)
#endif

// ---------------------------------------------------------------- the butterfly product
// w*y mod q for the butterflies, result in [0, GROWTH*q) for ANY word y.
//  32-bit words: exact Shoup, GROWTH = 2  (IMAD.HI + 2 IMAD).
//  64-bit words: GROWTH = 3.  On sm_100a IMAD.WIDE / IMAD.HI issue at half the IMAD.LO rate
//    (measured, tools/ubench), so the 64x64 high product is the expensive part.  The lowest partial
//    product y0*p0 is dropped: h' = floor((y1*p1*2^64 + (y0*p1 + y1*p0)*2^32) / 2^64) is h or h-1, which
//    costs one more q of range (the compile-time bound tracker in kernels.cuh pays for it with a
//    few extra top-bit reductions) and saves one of the four wide multiplies.
//  (A variant forming h*q with shifts for q = 2^60 - 2^14 + 1 was measured slower -- 10.1 vs 11.3 M
//   polymul/s -- because it trades 3 multiplies for ~8 ALU instructions and the kernel is issue-bound on
//   that side; see DESIGN.md.)
template <typename W> struct Growth;
template <> struct Growth<uint32_t> { static constexpr int G = 2; };
template <> struct Growth<uint64_t> { static constexpr int G = 3; };

TNTT_HD uint32_t shoup_lazy(uint32_t y, uint32_t w, uint32_t wp, const Mod<uint32_t> &m) {
    return shoup_mul(y, w, wp, m.nq);
}
TNTT_HD uint64_t shoup_lazy(uint64_t y, uint64_t w, uint64_t wp, const Mod<uint64_t> &m) {
#if defined(__CUDA_ARCH__)
    uint32_t y0, y1, w0, w1, p0, p1, h0, h1, lo, hi;
    unpack64(y, y0, y1);
    unpack64(w, w0, w1);
    unpack64(wp, p0, p1);
    asm("{\n\t.reg .u32 s0, s1, c;\n\t"
        "mul.lo.u32 s0, %2, %5;\n\t"
        "mul.hi.u32 s1, %2, %5;\n\t"
        "mad.lo.cc.u32 s0, %3, %4, s0;\n\t"
        "madc.hi.cc.u32 s1, %3, %4, s1;\n\t"
        "addc.u32 c, 0, 0;\n\t"
        "mad.lo.cc.u32 %0, %3, %5, s1;\n\t"
        "madc.hi.u32 %1, %3, %5, c;\n\t}"
        : "=r"(h0), "=r"(h1) : "r"(y0), "r"(y1), "r"(p0), "r"(p1));
    uint32_t n0, n1;
    unpack64(m.nq, n0, n1);
    uint64_t acc;   // low 64 bits of y*w + h*nq: one IMAD.WIDE chain plus four IMAD.LO into the high word
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(acc) : "r"(y0), "r"(w0));
#if defined(TNTT_X_TPART_REORDER)
    // everything that does not depend on the quotient first: only 3 multiplies follow h instead of 5
    unpack64(acc, lo, hi);
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(y0), "r"(w1));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(y1), "r"(w0));
    acc = pack64(lo, hi);
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(h0), "r"(n0));
    unpack64(acc, lo, hi);
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(h0), "r"(n1));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(h1), "r"(n0));
    return pack64(lo, hi);
#endif
#if defined(TNTT_X_SOLINAS_CROSS)
    // experiment, q = 2^60 - 2^14 + 1 only: nq = 2^64 - q has n0 = 2^14 - 1, n1 = -2^28 (mod 2^32), so the two
    // cross products h0*n1 + h1*n0 are (h1 << 14) - h1 - (h0 << 28): ALU work instead of multiplier work
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(h0), "r"(n0));
    unpack64(acc, lo, hi);
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(y0), "r"(w1));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(y1), "r"(w0));
    {
        uint32_t t1, t2;
#if TNTT_X_SOLINAS_CROSS == 1
        asm("shl.b32 %0, %1, 14;" : "=r"(t1) : "r"(h1));
        asm("shl.b32 %0, %1, 28;" : "=r"(t2) : "r"(h0));
        asm("sub.u32 %0, %0, %1;" : "+r"(t1) : "r"(h1));
        asm("{\n\t.reg .u32 z;\n\tsub.u32 z, %1, %2;\n\tadd.u32 %0, %0, z;\n\t}" : "+r"(hi) : "r"(t1), "r"(t2));
#elif TNTT_X_SOLINAS_CROSS == 2
        // funnel shifts cannot be expressed as IMAD: shf.l.wrap(lo = zero, hi = x, s) == x << s
        const uint32_t z = (uint32_t)m.zero;
        asm("shf.l.wrap.b32 %0, %1, %2, 14;" : "=r"(t1) : "r"(z), "r"(h1));
        asm("shf.l.wrap.b32 %0, %1, %2, 28;" : "=r"(t2) : "r"(z), "r"(h0));
        hi = hi + t1 - h1;          // three-input adds stay on the ALU
        hi = hi - t2 + z;
#endif
    }
    return pack64(lo, hi);
#endif
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(h0), "r"(n0));
    unpack64(acc, lo, hi);
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(h0), "r"(n1));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(h1), "r"(n0));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(y0), "r"(w1));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(y1), "r"(w0));
    return pack64(lo, hi);
#else
    const uint32_t y0 = (uint32_t)y, y1 = (uint32_t)(y >> 32), p0 = (uint32_t)wp, p1 = (uint32_t)(wp >> 32);
    const unsigned __int128 mid = (unsigned __int128)((uint64_t)y0 * p1) + (uint64_t)y1 * p0;
    const uint64_t h = (uint64_t)y1 * p1 + (uint64_t)(mid >> 32);
    return y * w - h * m.q;
#endif
}

// Cooley-Tukey butterfly on lazy values: (x, y) <- (x + w*y, x - w*y + G*q).  y may be ANY word (the
// product reduces it); the bound of both outputs is bound(x) + G*q.
template <typename W> TNTT_HD void ct_butterfly(W &x, W &y, const Tw<W> &t, const Mod<W> &m) {
    const W v = shoup_lazy(y, t.w, t.wp, m);
    TNTT_AUDIT_BUTTERFLY(x, v, m.qg);
    y = x - v + m.qg;
    if constexpr (sizeof(W) == 8) x = x + v + m.zero;
    else x = x + v;
}

// First stage of a decimation-in-time transform: its twiddle is 1, so no product is needed, only
// (x, y) <- (x + y, x - y + c) with c a multiple of q that is >= y (Mod::triv_c, or 2q in the Solinas kernels).
template <typename W> TNTT_HD void trivial_butterfly(W &x, W &y, W c) {
#if !defined(__CUDA_ARCH__) && defined(TNTT_AUDIT_RANGES)
    {
        const unsigned __int128 lim = (unsigned __int128)1 << WordTraits<W>::BITS;
        if ((unsigned __int128)x + y >= lim || y > c || (unsigned __int128)x + c >= lim + y) ++g_range_violations;
    }
#endif
    const W s = x + y;
    y = x - y + c;
    x = s;
}

// Compile-time range bookkeeping.  `red` selects how lazy values are kept below 2^BITS:
//   0: never (q is small enough for the whole transform, host::lazy_full_ok)
//   1: csub_top() -- test the top bit, subtract floor(2^(BITS-1)/q) q.  Bounds are counted in units of 2^(BITS-4)
//      (>= q because q < 2^(BITS-4)); a register below b units comes out below max(8, b - 7) units
//      (top_sub > 2^(BITS-1) - q > 7 units).
//   2: solinas_reduce() for q = 2^60 - 2^14 + 1 (64-bit words): x - (x >> 60) q < q + 2^18, a FULL reduction in three
//      instructions.  Bounds are counted in units of q (16 q < 2^64); a reduced register is below 2 units.
// Every register entering butterfly stage number `stage` (counted from the start of a transform whose inputs are
// below `b0` units) is below bound_at(...) units; a stage whose outputs could pass 16 units first reduces its
// un-multiplied inputs (the multiplied ones go through the product, which accepts any word).
//   3: no lazy values at all -- the reference's own arithmetic: Barrett products of canonical operands
//      (rtl/barrett_reduction.v:23-29) and fully reducing adds / subtracts (rtl/mod_add.v:14-15, mod_sub.v:15-17).
//      Kept as a measured alternative (profiles/r02_variant_sweep.jsonl), not as a default: see DESIGN.md.
TNTT_CX int reduce_result_bound(int red, int bound_in) { return red == 2 ? 2 : (bound_in - 7 > 8 ? bound_in - 7 : 8); }
#if defined(TNTT_X_NO_REDUCE)
TNTT_CX bool stage_needs_reduction(int, int, int) { return false; }   // what-if only: wrong results
#else
TNTT_CX bool stage_needs_reduction(int red, int g, int bound_in) { return (red == 1 || red == 2) && bound_in + g > 16; }
#endif
TNTT_CX int bound_after_stage(int red, int g, int bound_in) {
    if (red == 3) return 1;   // canonical in, canonical out
    return (stage_needs_reduction(red, g, bound_in) ? reduce_result_bound(red, bound_in) : bound_in) + g;
}
TNTT_CX int bound_at(int red, int g, int b0, int stage) {
    int b = b0;
    for (int s = 0; s < stage; ++s) b = bound_after_stage(red, g, b);
    return b;
}
// DIT transforms whose first stage is multiplication-free, (x, y) <- (x + y, x - y + c) with c a multiple of q >= y:
//   red 1: c = Mod::triv_c < 8 units serves inputs below b0 <= 7 units; outputs below b0 + 8
//   red 2: c = 2 q serves inputs below b0 <= 2 units (of q); outputs below 2 b0
constexpr int kTrivMaxIn = 7;
#if defined(TNTT_NO_TRIVIAL)
TNTT_CX bool dit_trivial_ok(int, int) { return false; }
#else
TNTT_CX bool dit_trivial_ok(int red, int b0) { return !red || red == 3 || (red == 2 ? b0 <= 2 : b0 <= kTrivMaxIn); }
#endif
TNTT_CX int dit_bound_at(int red, int g, int b0, int stage) {
    if (red == 3) return 1;
    if (!dit_trivial_ok(red, b0) || stage == 0) return bound_at(red, g, b0, stage);
    int b = red == 2 ? 2 * b0 : b0 + 8;
    for (int s = 1; s < stage; ++s) b = bound_after_stage(red, g, b);
    return b;
}

// RED 2 only (full reductions are three instructions there): the first DIT pass keeps index bits 0 .. LOGR-1 in the
// register index, so in EVERY stage of that pass the butterflies with twiddle index j = 0 multiply by root^0 = 1 and
// the choice is made at compile time -- (x, y) <- (x + y, x - y + c) with c = bound(y) q instead of a product.
// Such a butterfly doubles a bound instead of adding G, so this pass is tracked per register (units of q); an input
// is reduced first where its butterfly's outputs could pass 16 units.  Up to 4 + 2 + 1 products per thread fewer (R = 16).
#if defined(TNTT_NO_J0_TRIVIAL)
TNTT_CX bool dit2_j0_trivial() { return false; }
#else
TNTT_CX bool dit2_j0_trivial() { return true; }
#endif
// Measured on B200 (profiles/r02_whatif_j0.log, N = 4096 / 60-bit fused kernel): stages 0-1 (4 products fewer per
// thread) 13.53 -> 13.68 M polymul/s; stages 0-2 the same (2 products fewer, 2 reductions more); all four stages 13.61
// (3 reductions more and 8 bytes of spills at 128 registers).  Hence the shortcut stops after stage 1.
#if defined(TNTT_J0_MAXB)
constexpr int kJ0MaxStage = TNTT_J0_MAXB;   // experiments: only stages 0 .. TNTT_J0_MAXB take the shortcut
#else
constexpr int kJ0MaxStage = 1;
#endif
constexpr int kMaxR = 32;
struct RegBounds { int b[kMaxR]; };
struct Dit2Step { int bx, by; bool red_x, red_y; };   // bounds entering the butterfly after the reductions it needs
TNTT_CX Dit2Step dit2_step(bool trivial, int g, int bx, int by) {
    Dit2Step s{bx, by, false, false};
    if (trivial) {
        if (s.bx + s.by > 16) { s.by = 2; s.red_y = true; }
        if (s.bx + s.by > 16) { s.bx = 2; s.red_x = true; }
    } else if (s.bx + g > 16) { s.bx = 2; s.red_x = true; }
    return s;
}
// bounds of the 2^logr registers of a thread ENTERING stage `stage` (0 .. logr) of the first DIT pass, inputs below b0
TNTT_CX RegBounds dit2_pass0_bounds(int g, int b0, int logr, int stage) {
    RegBounds r{};
    for (int k = 0; k < (1 << logr); ++k) r.b[k] = b0;
    for (int B = 0; B < stage; ++B)
        for (int k0 = 0; k0 < (1 << logr); ++k0) {
            if (k0 & (1 << B)) continue;
            const int k1 = k0 | (1 << B);
            const bool trivial = (k0 & ((1 << B) - 1)) == 0 && B <= kJ0MaxStage;
            const Dit2Step s = dit2_step(trivial, g, r.b[k0], r.b[k1]);
            r.b[k0] = r.b[k1] = trivial ? s.bx + s.by : s.bx + g;
        }
    return r;
}
// uniform bound entering stage `stage` of a RED 2 DIT transform whose first pass holds `logr` stages
TNTT_CX int dit2_bound_at(int g, int b0, int logr, int stage) {
    const RegBounds r = dit2_pass0_bounds(g, b0, logr, stage < logr ? stage : logr);
    int b = 0;
    for (int k = 0; k < (1 << logr); ++k) b = r.b[k] > b ? r.b[k] : b;
    for (int s = logr; s < stage; ++s) b = bound_after_stage(2, g, b);
    return b;
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

// ---------------------------------------------------------------- q = 2^60 - 2^14 + 1 (the reference's 60-bit prime)
// 2^60 = 2^14 - 1 (mod q), so a value splits as A + B 2^60 = A + B (2^14 - 1): reductions are shifts and adds on the
// ALU pipe instead of products on the (saturated) multiplier pipe.  Selected at plan time (tntt_plan_info.solinas)
// for rtl/ntt_poly_mult.sv:16-26's modulus only; every other modulus keeps csub_top() / mont_mul().
constexpr uint64_t kSolinasQ = (1ull << 60) - (1ull << 14) + 1;
// any word x -> x - (x >> 60) q = (x mod 2^60) + (x >> 60)(2^14 - 1) < 2^60 + 15 * 2^14 < q + 2^18:
// LOP3 + SHF + one IMAD.WIDE (the product lands on the masked word pair)
TNTT_HD uint64_t solinas_reduce(uint64_t x) {
#if defined(__CUDA_ARCH__)
    uint32_t lo, hi;
    unpack64(x, lo, hi);
    const uint32_t t = hi >> 28;
    uint64_t r = pack64(lo, hi & 0x0FFFFFFFu);
    asm("mad.wide.u32 %0, %1, 16383, %0;" : "+l"(r) : "r"(t));
    return r;
#else
    return (x & ((1ull << 60) - 1)) + (x >> 60) * 16383ull;
#endif
}
// u * v mod q for ANY words u, v; result < 2^60 + 2^37 (below 2 q).  Four wide multiplies for the 128-bit product,
// then two folds: P = A + B 2^60 -> S = A + (B << 14) - B < 2^83 -> A' + B' 2^60 -> A' + (B' << 14) - B'.
TNTT_HD uint64_t solinas_mul(uint64_t u, uint64_t v) {
#if defined(__CUDA_ARCH__)
    uint32_t u0, u1, v0, v1;
    unpack64(u, u0, u1);
    unpack64(v, v0, v1);
    uint64_t t0, t1, t2, t3;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(t0) : "r"(u0), "r"(v0));
    uint32_t p0, t0h;
    unpack64(t0, p0, t0h);
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t1) : "r"(u0), "r"(v1), "l"((uint64_t)t0h));
    uint32_t t1l, t1h;
    unpack64(t1, t1l, t1h);
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t2) : "r"(u1), "r"(v0), "l"((uint64_t)t1l));
    uint32_t p1, t2h;
    unpack64(t2, p1, t2h);
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t3) : "r"(u1), "r"(v1), "l"((uint64_t)t1h));
    uint32_t p2, p3;
    unpack64(t3, p2, p3);
    asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, 0;" : "+r"(p2), "+r"(p3) : "r"(t2h));
    const uint32_t a1 = p1 & 0x0FFFFFFFu, b2 = p3 >> 28;
    uint32_t b0, b1, c1, c2;
    asm("shf.r.wrap.b32 %0, %1, %2, 28;" : "=r"(b0) : "r"(p1), "r"(p2));
    asm("shf.r.wrap.b32 %0, %1, %2, 28;" : "=r"(b1) : "r"(p2), "r"(p3));
    const uint32_t c0 = b0 << 14;
    asm("shf.l.wrap.b32 %0, %1, %2, 14;" : "=r"(c1) : "r"(b0), "r"(b1));
    asm("shf.l.wrap.b32 %0, %1, %2, 14;" : "=r"(c2) : "r"(b1), "r"(b2));
    uint32_t s0, s1, s2;
    asm("add.cc.u32 %0, %3, %5;\n\taddc.cc.u32 %1, %4, %6;\n\taddc.u32 %2, %7, 0;\n\t"
        "sub.cc.u32 %0, %0, %8;\n\tsubc.cc.u32 %1, %1, %9;\n\tsubc.u32 %2, %2, %10;"
        : "=&r"(s0), "=&r"(s1), "=&r"(s2) : "r"(p0), "r"(a1), "r"(c0), "r"(c1), "r"(c2), "r"(b0), "r"(b1), "r"(b2));
    const uint32_t a1b = s1 & 0x0FFFFFFFu;
    uint32_t bb;
    asm("shf.r.wrap.b32 %0, %1, %2, 28;" : "=r"(bb) : "r"(s1), "r"(s2));
    const uint32_t d0 = bb << 14, d1 = bb >> 18;
    uint32_t r0, r1;
    asm("add.cc.u32 %0, %2, %4;\n\taddc.u32 %1, %3, %5;\n\tsub.cc.u32 %0, %0, %6;\n\tsubc.u32 %1, %1, 0;"
        : "=&r"(r0), "=&r"(r1) : "r"(s0), "r"(a1b), "r"(d0), "r"(d1), "r"(bb));
    return pack64(r0, r1);
#else
    const unsigned __int128 P = (unsigned __int128)u * v;
    const uint64_t M = (1ull << 60) - 1;
    const unsigned __int128 B = P >> 60;
    const unsigned __int128 S = (P & M) + (B << 14) - B;
    const uint64_t B2 = (uint64_t)(S >> 60);
    return ((uint64_t)S & M) + (B2 << 14) - B2;
#endif
}
TNTT_HD uint32_t solinas_reduce(uint32_t x) { return x; }            // 32-bit words never use the Solinas path
TNTT_HD uint32_t solinas_mul(uint32_t u, uint32_t) { return u; }

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

// rtl/ntt_butterfly.v:43-72 literally: canonical (x, y) <- (x + w y mod q, x - w y mod q) with the Barrett product
template <typename W> TNTT_HD void ct_butterfly_barrett(W &x, W &y, const Tw<W> &t, const Mod<W> &m) {
    const W v = barrett_mul(y, t.w, m);
    y = csub((W)(x - v + m.q), m.q);     // rtl/mod_sub.v:15-17
    x = csub((W)(x + v), m.q);           // rtl/mod_add.v:14-15
}

}  // namespace tntt
