// Per-thread building blocks of the batched NTT kernels (sm_100a).
//
// How the reference's hardware dataflow is re-expressed:
//   rtl/ntt_cg_address_gen.v:57-117 + rtl/ntt_coeff_banks.v:159-320 (constant-geometry
//   addressing over ping-pong banks, bit-reversal on load)
//        -> a polynomial is held by P = N/R threads, R coefficients each, in registers.
//           log2(R) butterfly stages run register-to-register; between such "passes" the
//           coefficients are regrouped through an XOR-swizzled shared-memory tile
//           (elem()/spos() below), which is bank-conflict free for every access pattern
//           used (tests/test_layout.py enumerates them).
//   rtl/twiddle_bram_multiport.v:21-66 -> twiddle+Shoup pairs read through the read-only
//           path; tables are laid out so that a warp's request is contiguous or a broadcast.
//   rtl/ntt_control_parallel.v:58-138 -> fully unrolled stage loops; __syncthreads() is the
//           STAGE_DRAIN.
//   rtl/ntt_pointwise_mult.v:17-42 -> pointwise product on registers between the last forward
//           and first inverse pass (same thread<->coefficient layout), no memory traffic.
//
// Two transforms are used:
//   fwd_pass : merged-psi Cooley-Tukey, natural order in -> bit-reversed order out
//              (the psi^i twist of new_reference/cg_ntt.py:82-83 is folded into the twiddles)
//   dit_pass : cyclic decimation-in-time with omega^-1 (or any root), bit-reversed in -> natural out;
//              psi^-i * N^-1 (cg_ntt.py:74,91-92) is one multiply at the store.
// Everything here is __host__ __device__ and takes the thread id as an argument so that
// tests/host_emul.cpp can run the very same code on the CPU.
#pragma once
#if defined(TNTT_DEBUG_BOUNDS)
#include <cassert>
#endif
#include "modarith.cuh"

// how many twiddles (TG) / store-table entries (POST_GROUP) are fetched ahead of their use: Cfg::TG / Cfg::POST_GROUP
// (measured per shape on B200, profiles/r02_whatif_*.log); -DTNTT_TG / -DTNTT_POST_GROUP override them for experiments
#ifdef TNTT_TG
constexpr int kTgOverride = TNTT_TG;
#else
constexpr int kTgOverride = 0;
#endif
#ifdef TNTT_POST_GROUP
constexpr int kPostGroupOverride = TNTT_POST_GROUP;
#else
constexpr int kPostGroupOverride = 0;
#endif

namespace tntt {

TNTT_CX int cmax(int a, int b) { return a > b ? a : b; }
TNTT_CX int cmin(int a, int b) { return a < b ? a : b; }
TNTT_CX int cbitrev(int v, int bits) {   // compile-time bit reversal (register indices)
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((v >> i) & 1) << (bits - 1 - i);
    return r;
}

// Geometry of one kernel variant: W word, N = 2^LOGN coefficients, R = 2^LOGR per thread,
// PPC polynomials per CTA, NA operands transformed side by side (sharing twiddle loads).
// PAD_ = 1: the tile is padded by one word per R instead of XOR-swizzled.  Slots are then linear in the register index
// for every layout, so an exchange needs one base address per access pattern and immediate offsets (the swizzle costs
// a LOP3 + LEA per access).  Conflict-free for every pattern of the 64-bit shapes and of the 32-bit N = 256 / 1024
// shapes; the 32-bit N = 4096 shape pays one two-way conflict in the accesses with the register field at bit 8 and is
// still 5 % faster (tests/test_layout.py enumerates all of them; profiles/r02_whatif_u32.log).
template <typename W_, int LOGN_, int LOGR_, int PPC_, int PAD_ = 0> struct Cfg {
    using W = W_;
    static constexpr int PAD = PAD_;
    static constexpr int LOGN = LOGN_, LOGR = LOGR_, N = 1 << LOGN, R = 1 << LOGR;
    static constexpr int LOGP = LOGN - LOGR, P = 1 << LOGP;  // threads per polynomial
    static constexpr int PPC = PPC_, THREADS = P * PPC;
    static constexpr int NPASS = (LOGN + LOGR - 1) / LOGR;
    static constexpr int BANK_MASK = (1 << WordTraits<W>::BANK_BITS) - 1;
    // the padded shapes run two operands side by side at 128 registers: deeper twiddle groups, shallower store groups
    static constexpr bool WIDE_PAD = PAD_ && sizeof(W_) == 8;
    static constexpr int TG = kTgOverride ? kTgOverride : (WIDE_PAD ? 8 : 4);
    static constexpr int POST_GROUP = kPostGroupOverride ? kPostGroupOverride : (WIDE_PAD ? 2 : 4);
    // NA > 1: both operands' CTA-wide regroupings share one pair of barriers (measured: only pays in the padded shapes)
#ifdef TNTT_MERGE_EXCH
    static constexpr int MERGE_EXCH = TNTT_MERGE_EXCH;   // bit 0: the warp-local exchanges, bit 1: the CTA-wide ones
#else
    static constexpr int MERGE_EXCH = WIDE_PAD ? 2 : 0;
#endif
    // per-thread twiddle tables are prefetched one pass ahead only when they are too big to stay in L1
    #if defined(TNTT_FORCE_PREFETCH)
    static constexpr bool PREFETCH = true;
#elif defined(TNTT_X_NO_PREFETCH)
    static constexpr bool PREFETCH = false;   // what-if only
#else
    // (64 KB per table and more; the 32 KB tables of the 32-bit N = 4096 shape stay in L1 next to its three tiles, and
    // prefetching them costs 5 % there: 42.9 -> 45.3 M polymul/s, profiles/r02_whatif_u32.log)
    static constexpr bool PREFETCH = (size_t)N * 2 * sizeof(W) >= 65536;
#endif
    static_assert(LOGN >= LOGR, "a thread cannot hold more than the polynomial");
    // forward pass p works on index bits [fwd_lo(p), fwd_bhi(p)), high bits first
    static TNTT_CX int fwd_lo(int p) { return cmax(LOGN - (p + 1) * LOGR, 0); }
    static TNTT_CX int fwd_bhi(int p) { return LOGN - p * LOGR; }
    // inverse (DIT) pass p works on index bits [inv_blo(p), inv_bhi(p)), low bits first;
    // its register field starts at inv_lo(p)
    static TNTT_CX int inv_lo(int p) { return cmin(p * LOGR, LOGN - LOGR); }
    static TNTT_CX int inv_blo(int p) { return p * LOGR; }
    static TNTT_CX int inv_bhi(int p) { return cmin((p + 1) * LOGR, LOGN); }
    // coefficient index held in register k of thread tid when the register field sits at bit LO
    template <int LO> static TNTT_HD int elem(int tid, int k) {
        return ((tid >> LO) << (LO + LOGR)) | (k << LO) | (tid & ((1 << LO) - 1));
    }
    // shared-memory slot of CTA-wide element index E (= poly_in_cta * N + coefficient index)
    static constexpr int TILE = PPC * N + (PAD ? PPC * N / R : 0);   // words of one tile
    static TNTT_HD int spos(int E) {
        if constexpr (PAD) return E + (E >> LOGR);
        else return E ^ ((E >> LOGR) & BANK_MASK);
    }
    // size of the transposed last-forward-pass twiddle table
    static constexpr int FWD_LAST_ENTRIES = (R - 1) * P;
};

// Device view of the tables of one plan for one kernel variant.  The first MAX_R entries of each
// pyramid are also carried BY VALUE in the kernel parameters: the first forward pass and the first
// inverse pass use the same twiddles in every thread, and with compile-time indices into the
// parameter block they become constant-bank operands of the IMADs (no load, no register) -- the
// per-CTA "twiddle cache" of rtl/twiddle_bram_multiport.v for the uniform stages.
constexpr int MAX_R = 32;
template <typename W> struct DitTables {
    const Tw<W> *pyr;       // [N] DIT pyramid of a root: entry t+j = root^(j*N/(2t)), t = 2^b, j < t
    Tw<W> head[MAX_R];      // pyr[0 .. MAX_R)
};
template <typename W> struct PolymulTables {
    const Tw<W> *fwd_pyr;   // [N]  psi^bitrev pyramid: entry m+i, m = 2^stage
    const Tw<W> *fwd_last;  // [(R-1)*P] the same values for the last forward pass, [slot][tid]
    const Tw<W> *post;      // [N]  psi^-i * N^-1 * 2^BITS (the 2^BITS undoes the Montgomery pointwise product)
    Tw<W> fwd_head[MAX_R];  // fwd_pyr[0 .. MAX_R)
    DitTables<W> inv;       // omega^-1
};
template <typename W> struct TransformTables {
    DitTables<W> dit;   // pyramid of the root (omega or omega^-1)
    const Tw<W> *pre;   // [N] multiply on load (psi^i) or nullptr
    const Tw<W> *post;  // [N] multiply on store (psi^-i N^-1) or nullptr -> post_uniform
    Tw<W> post_uniform; // {1, floor(2^BITS/q)} (forward) or {N^-1, ...} (inverse)
    int reduce_input;   // inputs may be any word: reduce on load
};

#if defined(TNTT_X_NO_TW_LOADS) && defined(__CUDACC__)
__constant__ unsigned long long x_fake_tw[2] = {431606828070683274ull, 6905709249130932383ull};   // what-if only
#endif
template <typename W> TNTT_HD Tw<W> ld_tw(const Tw<W> *p) {
#if defined(__CUDA_ARCH__) && defined(TNTT_X_NO_TW_LOADS)
    return Tw<W>{(W)x_fake_tw[0], (W)x_fake_tw[1]};
#elif defined(__CUDA_ARCH__) && defined(TNTT_X_TW_L1HIT)
    // what-if only: every twiddle load hits the same 512 bytes (always in L1)
    p = reinterpret_cast<const Tw<W> *>((reinterpret_cast<unsigned long long>(p) & ~0xFFFFFull)) + (threadIdx.x & 31);
    const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(p));
    return Tw<W>{(W)v.x, (W)v.y};
#elif defined(__CUDA_ARCH__) && defined(TNTT_X_TW_SMEM)
    // what-if only: every twiddle comes from shared memory (contents are whatever the tile holds)
    extern __shared__ __align__(16) unsigned char x_smem[];
    const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(x_smem + ((reinterpret_cast<unsigned long long>(p) & 0x7FF0ull)));
    return Tw<W>{(W)v.x, (W)v.y};
#elif defined(__CUDA_ARCH__)
    if constexpr (sizeof(W) == 4) {
        const uint2 v = __ldg(reinterpret_cast<const uint2 *>(p));
        return Tw<W>{v.x, v.y};
    } else {
        const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(p));
        return Tw<W>{v.x, v.y};
    }
#else
    return *p;
#endif
}
template <typename W> TNTT_HD W ld_stream(const W *p) {
#if defined(__CUDA_ARCH__)
    W v;  // coefficients are touched once: do not let them displace the twiddle tables from L1
    if constexpr (sizeof(W) == 4) asm volatile("ld.global.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else asm volatile("ld.global.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
#else
    return *p;
#endif
}
// pull a twiddle line towards L1 without holding a register for it (no-op on the host)
TNTT_HD void prefetch_l1(const void *p) {
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}
template <typename W> TNTT_HD void st_stream(W *p, W v) {
#if defined(__CUDA_ARCH__)
    __stcs(p, v);
#else
    *p = v;
#endif
}

// Lazy reduction before a stage: only the registers that enter it as the un-multiplied ("x") input
// need it -- the other input goes through shoup_mul(), which accepts
// any word, and every later stage inherits bound(x) + 2q.  KB = register-index bit of that stage.
template <class C, int KB, int RED> TNTT_HD void reduce_top_x(typename C::W (&x)[C::R], const Mod<typename C::W> &mod) {
#pragma unroll
    for (int k = 0; k < C::R; ++k)
        if (!(k & (1 << KB))) {
            if constexpr (RED == 2) x[k] = solinas_reduce(x[k]);
            else x[k] = csub_top(x[k], mod.top_sub);
        }
}

// twiddle pair from a shared-memory copy of a table (filled by a TMA bulk copy, see polymul_kernel)
template <typename W> TNTT_HD Tw<W> ld_tw_shared(const Tw<W> *p) {
#if defined(__CUDA_ARCH__)
    if constexpr (sizeof(W) == 4) {
        const uint2 v = *reinterpret_cast<const uint2 *>(p);
        return Tw<W>{v.x, v.y};
    } else {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(p);
        return Tw<W>{v.x, v.y};
    }
#else
    return *p;
#endif
}

// ---------------------------------------------------------------------------------------------
// merged-psi Cooley-Tukey pass (natural -> bit-reversed), NA operands sharing each twiddle
// ---------------------------------------------------------------------------------------------
// one stage (index bit B) of a forward pass; B is a template parameter so that every loop bound
// below is a compile-time constant and the register arrays never fall into local memory
// STAB: non-null = shared-memory copy of fwd_last (only read when SMEM_TW)
// PRE: the twiddle of the pass's first stage was loaded before the tile exchange (pre_t), so its
// latency overlaps the barriers instead of following them
// ALLPRE: all R-1 twiddles of the pass were loaded beforehand into pre_t[slot] (fwd_load_pass_twiddles)
template <class C, int PASS, int NA, int RED, int B, bool SMEM_TW = false, bool PRE = false, bool ALLPRE = false>
TNTT_HD void fwd_stage(typename C::W (&x)[NA][C::R], int tid, const PolymulTables<typename C::W> &tb,
                       const Mod<typename C::W> &mod, const Tw<typename C::W> *stab = nullptr,
                       const Tw<typename C::W> *pre_t = nullptr) {
    using W = typename C::W;
    constexpr int LO = C::fwd_lo(PASS);
    constexpr int kb = B - LO;            // bit of the register index this stage pairs over
    constexpr int s = C::LOGN - 1 - B;    // stage number: 2^s blocks
    constexpr int NG = C::R >> (kb + 1), NJ = 1 << kb;
    if constexpr (stage_needs_reduction(RED, Growth<W>::G, bound_at(RED, Growth<W>::G, 1, s))) {
#pragma unroll
        for (int a = 0; a < NA; ++a) reduce_top_x<C, kb, RED>(x[a], mod);
    }
#if !defined(TNTT_X_NO_REDUCE)
    // with RED set the tracked bound (units of 2^(BITS-4), or of q for RED 2) must stay within the word after every stage
    static_assert(RED == 0 || bound_after_stage(RED, Growth<W>::G, bound_at(RED, Growth<W>::G, 1, s)) <= 16,
                  "forward stage: lazy values could pass 2^BITS");
#endif
    // twiddles are fetched TG at a time, ahead of their butterflies, so that their latencies overlap
    constexpr int TG = NG < C::TG ? NG : C::TG;
#pragma unroll
    for (int g0 = 0; g0 < NG; g0 += TG) {
        Tw<W> tw[TG];
#pragma unroll
        for (int gi = 0; gi < TG; ++gi) {
            const int g = g0 + gi;
            if constexpr (ALLPRE)
                tw[gi] = pre_t[(1 << (C::LOGR - 1 - kb)) - 1 + g];
            else if constexpr (PRE && B == C::fwd_bhi(PASS) - 1 && NG == 1)
                tw[gi] = *pre_t;
            else if constexpr (LO == C::LOGP && C::R <= MAX_R)  // first pass: same twiddle in every thread -> kernel parameter
                tw[gi] = tb.fwd_head[(1 << s) + g];
            else if constexpr (LO == 0 && SMEM_TW)  // last pass, table staged in shared memory by TMA
                tw[gi] = ld_tw_shared(&stab[((1 << (C::LOGR - 1 - kb)) - 1 + g) * C::P + tid]);
            else if constexpr (LO == 0)  // last pass: every thread has its own twiddles -> transposed table, coalesced
                tw[gi] = ld_tw(&tb.fwd_last[((1 << (C::LOGR - 1 - kb)) - 1 + g) * C::P + tid]);
            else                    // middle passes: shared by 2^LO consecutive threads -> broadcast
                tw[gi] = ld_tw(&tb.fwd_pyr[(1 << s) + ((tid >> LO) << (C::LOGR - 1 - kb)) + g]);
        }
#pragma unroll
        for (int gi = 0; gi < TG; ++gi) {
            const int g = g0 + gi;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int k0 = (g << (kb + 1)) | j, k1 = k0 | (1 << kb);
#pragma unroll
                for (int a = 0; a < NA; ++a) {
                    if constexpr (RED == 3) ct_butterfly_barrett(x[a][k0], x[a][k1], tw[gi], mod);
                    else ct_butterfly(x[a][k0], x[a][k1], tw[gi], mod);
                }
            }
        }
    }
    if constexpr (B > LO) fwd_stage<C, PASS, NA, RED, B - 1, SMEM_TW, PRE, ALLPRE>(x, tid, tb, mod, stab, pre_t);
}
// all twiddles of forward pass PASS (its stages pair over register bits kb = bhi-LO-1 .. 0; a pass that does not
// run all log2 R stages leaves the low slots unused), slot = 2^(LOGR-1-kb) - 1 + g
template <class C, int PASS> TNTT_HD void fwd_load_pass_twiddles(Tw<typename C::W> (&t)[C::R - 1], int tid,
                                                                 const PolymulTables<typename C::W> &tb) {
    constexpr int LO = C::fwd_lo(PASS);
#pragma unroll
    for (int kb = C::fwd_bhi(PASS) - LO - 1; kb >= 0; --kb) {
        const int s = C::LOGN - 1 - (LO + kb);
#pragma unroll
        for (int g = 0; g < (C::R >> (kb + 1)); ++g) {
            const int slot = (1 << (C::LOGR - 1 - kb)) - 1 + g;
            if constexpr (LO == C::LOGP && C::R <= MAX_R) t[slot] = tb.fwd_head[(1 << s) + g];
            else if constexpr (LO == 0) t[slot] = ld_tw(&tb.fwd_last[slot * C::P + tid]);
            else t[slot] = ld_tw(&tb.fwd_pyr[(1 << s) + ((tid >> LO) << (C::LOGR - 1 - kb)) + g]);
        }
    }
}
// the single twiddle of the first stage of forward pass PASS (valid when that pass runs all log2 R stages)
template <class C, int PASS> TNTT_HD Tw<typename C::W> fwd_first_twiddle(int tid, const PolymulTables<typename C::W> &tb) {
    constexpr int LO = C::fwd_lo(PASS), B = C::fwd_bhi(PASS) - 1, s = C::LOGN - 1 - B;
    if constexpr (LO == 0) return ld_tw(&tb.fwd_last[tid]);
    else return ld_tw(&tb.fwd_pyr[(1 << s) + (tid >> LO)]);
}
// The last forward pass, the last inverse pass and the final scaling read per-thread-distinct
// table entries (about N*16 B each, together more than L1 holds next to the tiles).  They are
// prefetched one pass ahead so that the L2 latency is paid while butterflies are running.
template <class C> TNTT_HD void prefetch_fwd_last(int tid, const PolymulTables<typename C::W> &tb) {
    constexpr int LAST = C::NPASS - 1;
    constexpr int first_slot = (1 << (C::LOGR - (C::fwd_bhi(LAST) - C::fwd_lo(LAST)))) - 1;
#pragma unroll
    for (int slot = first_slot; slot < C::R - 1; ++slot) prefetch_l1(&tb.fwd_last[slot * C::P + tid]);
}
template <class C> TNTT_HD void prefetch_dit_last(int tid, const Tw<typename C::W> *pyr) {
    constexpr int LAST = C::NPASS - 1, LO = C::inv_lo(LAST);
    if (LO == 0) return;
#pragma unroll
    for (int b = C::inv_blo(LAST); b < C::inv_bhi(LAST); ++b)
#pragma unroll
        for (int j = 0; j < (1 << (b - LO)); ++j) prefetch_l1(&pyr[(1 << b) + (j << LO) + (tid & ((1 << LO) - 1))]);
}
template <class C> TNTT_HD void prefetch_post(int tid, const Tw<typename C::W> *post) {
    if (!post) return;
#pragma unroll
    for (int k = 0; k < C::R; ++k) prefetch_l1(&post[(k << C::LOGP) + tid]);
}

template <class C, int PASS, int NA, int RED, bool SMEM_TW = false, bool PRE = false, bool ALLPRE = false>
TNTT_HD void fwd_pass(typename C::W (&x)[NA][C::R], int tid, const PolymulTables<typename C::W> &tb,
                      const Mod<typename C::W> &mod, const Tw<typename C::W> *stab = nullptr,
                      const Tw<typename C::W> *pre_t = nullptr) {
    fwd_stage<C, PASS, NA, RED, C::fwd_bhi(PASS) - 1, SMEM_TW, PRE, ALLPRE>(x, tid, tb, mod, stab, pre_t);
}
// bound (units of 2^(BITS-4)) of the spectrum a forward transform of canonical input leaves in registers
template <class C, int RED> TNTT_CX int fwd_out_bound() {
    return bound_at(RED, Growth<typename C::W>::G, 1, C::LOGN);
}
// ... and of the pointwise product of two such spectra: Montgomery (red 0/1; u top-reduced first) u*v/2^BITS + q,
// Solinas (red 2) below 2 q
template <class C, int RED> TNTT_CX int pointwise_out_bound() {
    if (RED == 3) return 1;
    if (RED == 2) return 2;
    constexpr int bf = fwd_out_bound<C, RED>();
    return ((bf > 8 ? 8 : bf) * bf + 15) / 16 + 1;
}
// rtl/ntt_pointwise_mult.v:17-42 on the registers of one thread: u, v = lazy spectra of a and b
template <class C, int RED> TNTT_HD typename C::W pointwise_product(typename C::W u, typename C::W v, const Mod<typename C::W> &mod) {
    if constexpr (RED == 2) return solinas_mul(u, v);
    else if constexpr (RED == 3) return barrett_mul(u, v, mod);   // rtl/ntt_pointwise_mult.v:17-42 as it is
    else {
        if (RED && fwd_out_bound<C, RED>() > 8) u = csub_top(u, mod.top_sub);  // u < 2^(BITS-1): no overflow
        return mont_mul(u, v, mod);   // the 2^-BITS is undone by the store table
    }
}

// ---------------------------------------------------------------------------------------------
// cyclic decimation-in-time pass (bit-reversed -> natural) over a root's pyramid table
// ---------------------------------------------------------------------------------------------
// the butterflies of one DIT stage (index bit B), twiddles fetched TG at a time ahead of their use
template <class C, int PASS, int RED, int B, bool SMEM_TW, bool PRE, bool ALLPRE>
TNTT_HD void dit_stage_products(typename C::W (&x)[C::R], int tid, const DitTables<typename C::W> &dt, const Mod<typename C::W> &mod,
                                const Tw<typename C::W> *stab, const Tw<typename C::W> *pre_t) {
    using W = typename C::W;
    constexpr int LO = C::inv_lo(PASS);
    constexpr int kb = B - LO;
    constexpr int NG = C::R >> (kb + 1), NJ = 1 << kb;
    constexpr int TG = NJ < C::TG ? NJ : C::TG;
#pragma unroll
    for (int j0 = 0; j0 < NJ; j0 += TG) {
        Tw<W> tw[TG];
#pragma unroll
        for (int ji = 0; ji < TG; ++ji) {
            const int j = j0 + ji;
            if constexpr (ALLPRE) tw[ji] = pre_t[(1 << kb) - 1 + j];
            else if constexpr (PRE && B == C::inv_blo(PASS) && NJ == 1) tw[ji] = *pre_t;
            else if constexpr (LO == 0 && C::R <= MAX_R) tw[ji] = dt.head[(1 << B) + j];  // first pass: uniform -> kernel parameter
            else if constexpr (SMEM_TW && PASS + 1 == C::NPASS)   // last pass, pyr[2^blo ..) staged in shared memory by TMA
                tw[ji] = ld_tw_shared(&stab[(1 << B) - (1 << C::inv_blo(PASS)) + (j << LO) + (tid & ((1 << LO) - 1))]);
            else tw[ji] = ld_tw(&dt.pyr[(1 << B) + (j << LO) + (tid & ((1 << LO) - 1))]);
        }
#pragma unroll
        for (int ji = 0; ji < TG; ++ji) {
            const int j = j0 + ji;
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const int k0 = (g << (kb + 1)) | j, k1 = k0 | (1 << kb);
                if constexpr (RED == 3) ct_butterfly_barrett(x[k0], x[k1], tw[ji], mod);
                else ct_butterfly(x[k0], x[k1], tw[ji], mod);
            }
        }
    }
}

// IN_BND: bound of the transform's input in units of 2^(BITS-4) (only used when RED)
template <class C, int PASS, int RED, int IN_BND, int B, bool SMEM_TW = false, bool PRE = false, bool ALLPRE = false>
TNTT_HD void dit_stage(typename C::W (&x)[C::R], int tid, const DitTables<typename C::W> &dt, const Mod<typename C::W> &mod,
                       const Tw<typename C::W> *stab = nullptr, const Tw<typename C::W> *pre_t = nullptr) {
    using W = typename C::W;
    constexpr int LO = C::inv_lo(PASS);
    constexpr int kb = B - LO;
    constexpr int NG = C::R >> (kb + 1), NJ = 1 << kb;
    constexpr bool J0 = RED == 2 && dit2_j0_trivial() && C::LOGR <= 5;   // see dit2_pass0_bounds (modarith.cuh)
    if constexpr (J0 && PASS == 0) {
        // Solinas kernels, first pass: per-register bounds, the j = 0 butterflies of every stage are multiplication-free
        constexpr int G = Growth<W>::G;
        constexpr RegBounds rb = dit2_pass0_bounds(G, IN_BND, C::LOGR, B);
        static_assert(dit2_bound_at(G, IN_BND, C::LOGR, B + 1) <= 16, "first inverse pass: lazy values could pass 2^BITS");
        static_assert(C::R <= MAX_R && C::R <= kMaxR, "the first pass takes its twiddles from the kernel parameters");
#pragma unroll
        for (int g = 0; g < NG && B <= kJ0MaxStage; ++g) {   // j = 0: twiddle root^0 = 1
            const int k0 = g << (kb + 1), k1 = k0 | (1 << kb);
            const Dit2Step st = dit2_step(true, G, rb.b[k0], rb.b[k1]);
            if (st.red_y) x[k1] = solinas_reduce(x[k1]);
            if (st.red_x) x[k0] = solinas_reduce(x[k0]);
            trivial_butterfly(x[k0], x[k1], (W)(mod.q * (W)st.by));
        }
        constexpr int TG = NJ < C::TG ? NJ : C::TG;
#pragma unroll
        for (int j0 = 0; j0 < NJ; j0 += TG) {
            Tw<W> tw[TG];
#pragma unroll
            for (int ji = 0; ji < TG; ++ji)
                if (j0 + ji || B > kJ0MaxStage) tw[ji] = ALLPRE ? pre_t[(1 << kb) - 1 + j0 + ji] : dt.head[(1 << B) + j0 + ji];
#pragma unroll
            for (int ji = 0; ji < TG; ++ji) {
                const int j = j0 + ji;
                if (!j && B <= kJ0MaxStage) continue;
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const int k0 = (g << (kb + 1)) | j, k1 = k0 | (1 << kb);
                    if (dit2_step(false, G, rb.b[k0], rb.b[k1]).red_x) x[k0] = solinas_reduce(x[k0]);
                    ct_butterfly(x[k0], x[k1], tw[ji], mod);
                }
            }
        }
    } else if constexpr (J0) {
        // later passes: one bound for all registers, continued from the largest one the first pass leaves
        constexpr int G = Growth<W>::G, bin = dit2_bound_at(G, IN_BND, C::LOGR, B);
        if constexpr (stage_needs_reduction(RED, G, bin)) reduce_top_x<C, kb, RED>(x, mod);
        static_assert(bound_after_stage(RED, G, bin) <= 16, "inverse stage: lazy values could pass 2^BITS");
        dit_stage_products<C, PASS, RED, B, SMEM_TW, PRE, ALLPRE>(x, tid, dt, mod, stab, pre_t);
    } else if constexpr (B == 0 && dit_trivial_ok(RED, IN_BND)) {   // twiddle 1: no product at all
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            if constexpr (RED == 3) {
                trivial_butterfly(x[2 * g], x[2 * g + 1], mod.q);
                x[2 * g] = csub(x[2 * g], mod.q);
                x[2 * g + 1] = csub(x[2 * g + 1], mod.q);
            } else {
                trivial_butterfly(x[2 * g], x[2 * g + 1], RED == 2 ? mod.q2 : mod.triv_c);
            }
        }
    } else {
    if constexpr (stage_needs_reduction(RED, Growth<W>::G, dit_bound_at(RED, Growth<W>::G, IN_BND, B)))
        reduce_top_x<C, kb, RED>(x, mod);
#if !defined(TNTT_X_NO_REDUCE)
    static_assert(RED == 0 || bound_after_stage(RED, Growth<W>::G, dit_bound_at(RED, Growth<W>::G, IN_BND, B)) <= 16,
                  "inverse stage: lazy values could pass 2^BITS");
#endif
    dit_stage_products<C, PASS, RED, B, SMEM_TW, PRE, ALLPRE>(x, tid, dt, mod, stab, pre_t);
    }
    if constexpr (B + 1 < C::inv_bhi(PASS)) dit_stage<C, PASS, RED, IN_BND, B + 1, SMEM_TW, PRE, ALLPRE>(x, tid, dt, mod, stab, pre_t);
}
template <class C, int PASS, int RED, int IN_BND, bool SMEM_TW = false, bool PRE = false, bool ALLPRE = false>
TNTT_HD void dit_pass(typename C::W (&x)[C::R], int tid, const DitTables<typename C::W> &dt, const Mod<typename C::W> &mod,
                      const Tw<typename C::W> *stab = nullptr, const Tw<typename C::W> *pre_t = nullptr) {
    dit_stage<C, PASS, RED, IN_BND, C::inv_blo(PASS), SMEM_TW, PRE, ALLPRE>(x, tid, dt, mod, stab, pre_t);
}
// all twiddles of inverse pass PASS (stages B = blo .. bhi-1, kb = B - LO), slot = 2^kb - 1 + j; slot 0 of pass 0 is the trivial twiddle 1
template <class C, int PASS> TNTT_HD void dit_load_pass_twiddles(Tw<typename C::W> (&t)[C::R - 1], int tid,
                                                                 const DitTables<typename C::W> &dt) {
    constexpr int LO = C::inv_lo(PASS);
#pragma unroll
    for (int kb = C::inv_blo(PASS) - LO; kb < C::inv_bhi(PASS) - LO; ++kb) {
        const int B = LO + kb;
#pragma unroll
        for (int j = 0; j < (1 << kb); ++j) {
            const int slot = (1 << kb) - 1 + j;
            if constexpr (LO == 0 && C::R <= MAX_R) t[slot] = dt.head[(1 << B) + j];
            else t[slot] = ld_tw(&dt.pyr[(1 << B) + (j << LO) + (tid & ((1 << LO) - 1))]);
        }
    }
}
// the single twiddle of the first stage of inverse pass PASS (valid when its register field starts at that bit)
template <class C, int PASS> TNTT_HD Tw<typename C::W> dit_first_twiddle(int tid, const DitTables<typename C::W> &dt) {
    constexpr int LO = C::inv_lo(PASS), B = C::inv_blo(PASS);
    return ld_tw(&dt.pyr[(1 << B) + (tid & ((1 << LO) - 1))]);
}

// registers -> swizzled tile (layout with the register field at LO)
template <class C, int LO> TNTT_HD void tile_write(const typename C::W (&x)[C::R], typename C::W *tile, int pl, int tid) {
#pragma unroll
    for (int k = 0; k < C::R; ++k) tile[C::spos(pl * C::N + C::template elem<LO>(tid, k))] = x[k];
}
template <class C, int LO> TNTT_HD void tile_read(typename C::W (&x)[C::R], const typename C::W *tile, int pl, int tid) {
#pragma unroll
    for (int k = 0; k < C::R; ++k) x[k] = tile[C::spos(pl * C::N + C::template elem<LO>(tid, k))];
}

TNTT_HD int bitrev_n(int v, int bits) {
#if defined(__CUDA_ARCH__)
    return (int)(__brev((unsigned)v) >> (32 - bits));
#else
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((v >> i) & 1) << (bits - 1 - i);
    return r;
#endif
}

// coalesced row access: register k of thread tid <-> coefficient k*P + tid (field at LOGP)
template <class C> TNTT_HD void row_load(typename C::W (&x)[C::R], const typename C::W *row, int tid, bool active) {
#pragma unroll
    for (int k = 0; k < C::R; ++k) x[k] = active ? ld_stream(row + (k << C::LOGP) + tid) : (typename C::W)0;
}

// final multiply (psi^-i N^-1 ...) + canonical reduction + coalesced store.
// TABLE: 1 = per-coefficient table `post` (loaded GROUP entries at a time so that their L2 latencies
// overlap instead of one exposed load per coefficient), 0 = one uniform factor, -1 = decided at run time.
// (RED is accepted for symmetry with the other building blocks.  A Solinas form of this product for RED 2 -- four wide
// multiplies and ALU folds instead of the exact Shoup product's six wide + four narrow -- was measured SLOWER on B200,
// 13.05 vs 13.54 M polymul/s, profiles/r02_whatif_store.log: at the tail of the kernel nothing overlaps the ALU folds.)
template <class C, int TABLE = -1, int GROUP_ = C::POST_GROUP, int RED = 0>
TNTT_HD void row_store_scaled(const typename C::W (&x)[C::R], typename C::W *row, int tid, bool active,
                              const Tw<typename C::W> *post, const Tw<typename C::W> &post_uniform,
                              const Mod<typename C::W> &mod) {
    using W = typename C::W;
    constexpr int GROUP = GROUP_ < C::R ? GROUP_ : C::R;
    if (TABLE == 1 || (TABLE == -1 && post)) {
#pragma unroll
        for (int k0 = 0; k0 < C::R; k0 += GROUP) {
            Tw<W> t[GROUP];
#pragma unroll
            for (int j = 0; j < GROUP; ++j) t[j] = ld_tw(&post[((k0 + j) << C::LOGP) + tid]);
#pragma unroll
            for (int j = 0; j < GROUP; ++j) {
#if defined(TNTT_X_NO_POST)
                const W v = x[k0 + j] ^ t[j].w;   // what-if only
#else
                const W v = csub(shoup_mul(x[k0 + j], t[j].w, t[j].wp, mod.nq), mod.q);
#endif
                if (active) st_stream(row + ((k0 + j) << C::LOGP) + tid, v);
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < C::R; ++k) {
            const W v = csub(shoup_mul(x[k], post_uniform.w, post_uniform.wp, mod.nq), mod.q);
            if (active) st_stream(row + (k << C::LOGP) + tid, v);
        }
    }
}

constexpr int kTwBufBytes = 65536;   // shared twiddle buffer of the TMA variants (+ 16 B for the mbarrier)

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------------------------
// regroup registers through the shared tile: layout LO_FROM -> LO_TO
// ---------------------------------------------------------------------------------------------
// A polynomial's region of the tile is touched only by the P threads that own the polynomial.  When those sit
// inside one warp (N = 1024 with 32 coefficients per thread, N = 256), a warp barrier orders the exchange and
// the other warps of the CTA are not held up.
template <class C> __device__ __forceinline__ void tile_sync() {
    if constexpr (C::P <= 32) __syncwarp();
    else __syncthreads();
}
// A regrouping between the layouts LO_A and LO_B with min = 0 only moves coefficients among the 2^max aligned,
// consecutive threads that share tid >> max (they hold the same 2^(max + LOGR) consecutive coefficients before and
// after), and nobody else touches that part of the tile in either layout.  When such a group sits inside one warp
// the exchange needs a warp barrier only, and the warps of the CTA are free to drift apart (test_layout.py
// enumerates the thread sets).
template <class C, int LO_A, int LO_B> TNTT_CX bool exchange_is_warp_local() {
#if defined(TNTT_X_CTA_SYNC_ONLY)
    return C::P <= 32;
#else
    return C::P <= 32 || ((LO_A < LO_B ? LO_A : LO_B) == 0 && (LO_A < LO_B ? LO_B : LO_A) <= 5);
#endif
}
template <class C, int LO_A, int LO_B> __device__ __forceinline__ void exchange_sync() {
    if constexpr (exchange_is_warp_local<C, LO_A, LO_B>()) __syncwarp();
    else __syncthreads();
}
template <class C, int LO_FROM, int LO_TO>
__device__ __forceinline__ void exchange(typename C::W (&x)[C::R], typename C::W *tile, int pl, int tid) {
#if defined(TNTT_X_NO_EXCHANGE)
    return;   // what-if only: wrong results
#endif
    exchange_sync<C, LO_FROM, LO_TO>();  // everybody is done reading the tile's previous contents
    tile_write<C, LO_FROM>(x, tile, pl, tid);
    exchange_sync<C, LO_FROM, LO_TO>();
    tile_read<C, LO_TO>(x, tile, pl, tid);
}

// ---------------------------------------------------------------------------------------------
// TMA staging of the per-thread twiddle tables (TMA = 1 variants)
//
// The last forward pass, the last inverse pass and the final scaling read table entries that differ
// per thread: 3 x 60 KB + 64 KB per polymul, more than L1 keeps next to the tiles, so plain loads pay
// the L2 latency (ncu: long_scoreboard is a top stall).  Instead one thread issues a bulk async copy
// (cp.async.bulk, the TMA engine; SASS UBLKCP) of the whole table into a 64 KB shared buffer while the
// CTA computes the passes before it, and the CTA waits on an mbarrier just before the first use.
// ---------------------------------------------------------------------------------------------
struct TmaStage {
    unsigned bar;     // shared address of the mbarrier
    unsigned dst;     // shared address of the 64 KB table buffer
    unsigned phase;   // parity of the next completion to wait for
    __device__ __forceinline__ void init(void *bar_ptr, void *buf_ptr) {
        bar = (unsigned)__cvta_generic_to_shared(bar_ptr);
        dst = (unsigned)__cvta_generic_to_shared(buf_ptr);
        phase = 0;
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    // thread 0 only; every thread must already be past its last read of the buffer (a barrier)
    __device__ __forceinline__ void issue(const void *src, unsigned bytes) const {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
    }
    __device__ __forceinline__ void wait() {
        asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                     "@!p bra WAIT_%=;\n\t}" ::"r"(bar), "r"(phase) : "memory");
        phase ^= 1u;
    }
};

template <class C, int NA, int RED, bool TMA, int PASS = 0>
__device__ __forceinline__ void forward_all(typename C::W (&x)[NA][C::R], typename C::W *tile, int pl, int tid,
                                            const PolymulTables<typename C::W> &tb, const Mod<typename C::W> &mod,
                                            TmaStage *tma = nullptr, const Tw<typename C::W> *stab = nullptr,
                                            bool wait_table = false) {
    if constexpr (PASS < C::NPASS) {
        // first-stage twiddle of this pass: issue the load before the barriers of the exchange
        constexpr bool PRE = PASS > 0 && (C::fwd_bhi(PASS) - C::fwd_lo(PASS) == C::LOGR) && !(TMA && PASS + 1 == C::NPASS);
        Tw<typename C::W> t0{};
        if constexpr (PRE) t0 = fwd_first_twiddle<C, PASS>(tid, tb);
        if constexpr (PASS > 0) {
#if defined(TNTT_X_NO_EXCHANGE)
            if constexpr (false) {
#else
            if constexpr (NA > 1 && ((exchange_is_warp_local<C, C::fwd_lo(PASS - 1), C::fwd_lo(PASS)>() ? 1 : 2) & C::MERGE_EXCH)) {
                // one pair of barriers serves all operands (each has its own tile)
#endif
                exchange_sync<C, C::fwd_lo(PASS - 1), C::fwd_lo(PASS)>();
#pragma unroll
                for (int a = 0; a < NA; ++a) tile_write<C, C::fwd_lo(PASS - 1)>(x[a], tile + a * C::TILE, pl, tid);
                exchange_sync<C, C::fwd_lo(PASS - 1), C::fwd_lo(PASS)>();
#pragma unroll
                for (int a = 0; a < NA; ++a) tile_read<C, C::fwd_lo(PASS)>(x[a], tile + a * C::TILE, pl, tid);
            } else {
#pragma unroll
                for (int a = 0; a < NA; ++a) exchange<C, C::fwd_lo(PASS - 1), C::fwd_lo(PASS)>(x[a], tile + a * C::TILE, pl, tid);
            }
        }
#ifndef TNTT_PF_MASK
#define TNTT_PF_MASK 7     // which per-thread tables are prefetched one pass ahead: 1 = fwd_last, 2 = last inverse pass, 4 = store table
#endif
        if constexpr (PASS + 2 == C::NPASS && C::PREFETCH && !TMA && (TNTT_PF_MASK & 1)) prefetch_fwd_last<C>(tid, tb);
        if constexpr (TMA && PASS + 1 == C::NPASS) {
            if (wait_table) tma->wait();
        }
        fwd_pass<C, PASS, NA, RED, TMA, PRE>(x, tid, tb, mod, stab, &t0);
        forward_all<C, NA, RED, TMA, PASS + 1>(x, tile, pl, tid, tb, mod, tma, stab, wait_table);
    }
}

// PF: prefetch the last pass's twiddles and the store table one pass ahead
template <class C, int RED, int IN_BND, bool TMA, bool PF, int PASS = 0>
__device__ __forceinline__ void dit_all(typename C::W (&x)[C::R], typename C::W *tile, int pl, int tid,
                                        const DitTables<typename C::W> &dt, const Tw<typename C::W> *post,
                                        const Mod<typename C::W> &mod, TmaStage *tma = nullptr,
                                        const Tw<typename C::W> *stab = nullptr) {
    if constexpr (PASS < C::NPASS) {
#ifndef TNTT_INV_ALLPRE
#define TNTT_INV_ALLPRE 0
#endif
        // APRE: every twiddle of the pass is fetched before the exchange (the inverse of a kernel built for two operands
        // side by side has the registers of the second operand to spare)
        constexpr bool APRE = TNTT_INV_ALLPRE && PASS > 0 && !TMA && PF;
        constexpr bool PRE = !APRE && PASS > 0 && C::inv_lo(PASS) == C::inv_blo(PASS) && !(TMA && PASS + 1 == C::NPASS);
        Tw<typename C::W> t0{};
        Tw<typename C::W> tall[APRE ? C::R - 1 : 1];
        if constexpr (PRE) t0 = dit_first_twiddle<C, PASS>(tid, dt);
        if constexpr (APRE) dit_load_pass_twiddles<C, PASS>(tall, tid, dt);
#if defined(TNTT_X_NO_EXCHANGE)
        if constexpr (false) {
#else
        if constexpr (PASS > 0) {
#endif
            if constexpr (TMA) __syncthreads();  // everybody is done reading the tile and, for PASS 1, the forward twiddle buffer
            else exchange_sync<C, C::inv_lo(PASS - 1), C::inv_lo(PASS)>();
            if constexpr (TMA && PASS == 1) {
                constexpr int first = 1 << C::inv_blo(C::NPASS - 1);
                if (threadIdx.x == 0)
                    tma->issue(dt.pyr + first, (unsigned)((C::N - first) * sizeof(Tw<typename C::W>)));
            }
            tile_write<C, C::inv_lo(PASS - 1)>(x, tile, pl, tid);
            if constexpr (TMA) __syncthreads();
            else exchange_sync<C, C::inv_lo(PASS - 1), C::inv_lo(PASS)>();
            tile_read<C, C::inv_lo(PASS)>(x, tile, pl, tid);
        }
        if constexpr (PASS + 2 == C::NPASS && PF && !TMA && (TNTT_PF_MASK & 2)) prefetch_dit_last<C>(tid, dt.pyr);
        if constexpr (PASS + 1 == C::NPASS && PF && (TNTT_PF_MASK & 4)) prefetch_post<C>(tid, post);
        if constexpr (TMA && PASS + 1 == C::NPASS) tma->wait();
        if constexpr (APRE) dit_pass<C, PASS, RED, IN_BND, false, false, true>(x, tid, dt, mod, nullptr, tall);
        else dit_pass<C, PASS, RED, IN_BND, TMA, PRE>(x, tid, dt, mod, stab, &t0);
        dit_all<C, RED, IN_BND, TMA, PF, PASS + 1>(x, tile, pl, tid, dt, post, mod, tma, stab);
    }
}

// ---------------------------------------------------------------------------------------------
// THE hot path: c = a * b in Z_q[x]/(x^N+1), one HBM round trip per polynomial
// (new_reference/cg_ntt.py:78-92; rtl/ntt_poly_mult.sv:479-530 without its 20k copy cycles)
//   NA = 1: forward(a), park it in registers, forward(b)       (one tile)
//   NA = 2: forward(a) and forward(b) side by side, one twiddle load serves both (two tiles)
// ---------------------------------------------------------------------------------------------
//   STASH = 1 (NA = 1 only): a's spectrum waits in a second shared tile instead of in registers while
//              b is transformed (thread-private slots, [k][thread] order: no conflicts, no barrier)
//   TMA = 1: the per-thread twiddle tables are staged in shared memory by bulk async copies (see TmaStage)
template <class C, int NA, int RED, int MINB, int STASH = 0, int TMA = 0>
__global__ void __launch_bounds__(C::THREADS, MINB)
polymul_kernel(const typename C::W *a, const typename C::W *b, typename C::W *c,   // c may alias a or b: no __restrict__
               size_t batch, const __grid_constant__ PolymulTables<typename C::W> tb,
               const __grid_constant__ Mod<typename C::W> mod) {
#include "polymul_body.inc"
}

// ---------------------------------------------------------------------------------------------
// Multi-modulus (RNS) batches, SURVEY.md section 8 f3 (reports/final-report.tex:1811-1817): operands [L][batch][N],
// limb l holding the residues mod q_l.  ONE launch serves all limbs: blockIdx.y selects the limb's tables and
// modulus constants from an array carried BY VALUE in the kernel parameters, so the uniform twiddles of the first
// passes and the modulus constants stay constant-bank / uniform-register operands exactly as in the single-modulus
// kernel (an index into the parameter block costs a uniform add).  16 limbs of 64-bit tables are 18 KB of the 32 KB
// parameter space of sm_70+ under CUDA >= 12.1; more limbs take more launches (capi: tntt_rns_polymul).
// ---------------------------------------------------------------------------------------------
constexpr int kRnsMaxLimbs = 16;
template <typename W> struct RnsLimb {
    PolymulTables<W> tb;
    Mod<W> mod;
    const Tw<W> *untwist;   // [N] psi^-i N^-1 (tb.post carries the extra 2^BITS of the Montgomery pointwise product)
};
template <typename W> struct RnsLimbs { RnsLimb<W> limb[kRnsMaxLimbs]; };

template <class C, int NA, int RED, int MINB, int STASH = 0>
__global__ void __launch_bounds__(C::THREADS, MINB)
polymul_rns_kernel(const typename C::W *a, const typename C::W *b, typename C::W *c, size_t batch,
                   const __grid_constant__ RnsLimbs<typename C::W> limbs) {
    constexpr int TMA = 0;
    const PolymulTables<typename C::W> &tb = limbs.limb[blockIdx.y].tb;
    const Mod<typename C::W> &mod = limbs.limb[blockIdx.y].mod;
    const size_t limb_off = (size_t)blockIdx.y * batch * C::N;
    a += limb_off;
    b += limb_off;
    c += limb_off;
#include "polymul_body.inc"
}

// ---------------------------------------------------------------------------------------------
// Operands kept in the transform domain (one forward transform per operand, reused
// across many products -- the RLWE use of the reference, reports/final-report.tex:571-610).
//
// A "spectrum" row is the negacyclic NTT of a polynomial in the order the fused kernel holds it in registers
// between its last forward and first inverse pass, stored so that the access is the coalesced row pattern:
// word k*P + t of the row = bit-reversed-order element t*R + k (R, P of the plan's spectrum shape).  It is
// canonical ([0, q)), so tntt_pointwise applies to it directly (the product is order-agnostic).  Producing and
// consuming spectra needs no bit reversal, no twist pass and no regrouping beyond the fused kernel's own:
//   spectrum_forward_kernel  = the forward half of polymul_kernel          (ntt(twist(a)), cg_ntt.py:82-87)
//   spectrum_inverse_kernel  = its inverse half                            (untwist(cg_intt(.)), cg_ntt.py:90-92)
//   polymul_spectrum_kernel  = polymul_kernel with b's transform replaced by a load (b may be one shared row)
// ---------------------------------------------------------------------------------------------
// NATURAL = true turns the two kernels into natural-order transforms (cg_ntt / cg_intt and their twisted forms,
// new_reference/cg_ntt.py:29-75): position p = t*R + k of the bit-reversed-order spectrum is natural index
// bitrev(p) = (bitrev_R(k) << log2 P) | bitrev_P(t).  For P <= 32 a warp's permuted accesses still cover whole
// contiguous segments, so the permutation costs nothing; otherwise it goes through the tile once.  The tables
// decide the transform: merged-psi pyramid = ntt(twist(.)), cyclic pyramid (host::fwd_pyramid_cyclic) = cg_ntt.
template <class C, int RED, int MINB, bool NATURAL = false>
__global__ void __launch_bounds__(C::THREADS, MINB)
spectrum_forward_kernel(const typename C::W *in, typename C::W *out, size_t batch,
                        const __grid_constant__ PolymulTables<typename C::W> tb,
                        const __grid_constant__ Mod<typename C::W> mod) {
    using W = typename C::W;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    W *tile = reinterpret_cast<W *>(smem_raw);
    const int tid = threadIdx.x & (C::P - 1), pl = threadIdx.x >> C::LOGP;
    const size_t poly = (size_t)blockIdx.x * C::PPC + pl;
    const bool active = poly < batch;
    const size_t off = active ? poly * C::N : 0;
    W x[1][C::R];
    row_load<C>(x[0], in + off, tid, active);
    forward_all<C, 1, RED, false>(x, tile, pl, tid, tb, mod);
    auto canon = [&](W v) { return csub(shoup_mul(v, (W)1, mod.one_p, mod.nq), mod.q); };   // any word -> [0, q)
    if constexpr (!NATURAL) {
#pragma unroll
        for (int k = 0; k < C::R; ++k) {
            const W v = canon(x[0][k]);
            if (active) st_stream(out + off + (k << C::LOGP) + tid, v);
        }
    } else if constexpr (C::P <= 32) {
        const int bt = bitrev_n(tid, C::LOGP);
#pragma unroll
        for (int k = 0; k < C::R; ++k) {
            const W v = canon(x[0][k]);
            if (active) st_stream(out + off + (cbitrev(k, C::LOGR) << C::LOGP) + bt, v);
        }
    } else {
        const int bt = bitrev_n(tid, C::LOGP);
        tile_sync<C>();
#pragma unroll
        for (int k = 0; k < C::R; ++k) tile[C::spos(pl * C::N + ((cbitrev(k, C::LOGR) << C::LOGP) | bt))] = canon(x[0][k]);
        tile_sync<C>();
#pragma unroll
        for (int k = 0; k < C::R; ++k)
            if (active) st_stream(out + off + (k << C::LOGP) + tid, tile[C::spos(pl * C::N + (k << C::LOGP) + tid)]);
    }
}

// TABLE: the store multiplies by the per-coefficient table `post` (psi^-i N^-1: twisted inverse) or, TABLE = false,
// by the one factor `post_uniform` (N^-1: cg_intt)
template <class C, int RED, int MINB, bool NATURAL = false, bool TABLE = true>
__global__ void __launch_bounds__(C::THREADS, MINB)
spectrum_inverse_kernel(const typename C::W *in, typename C::W *out, size_t batch,
                        const __grid_constant__ DitTables<typename C::W> inv, const Tw<typename C::W> *__restrict__ post,
                        const Tw<typename C::W> post_uniform, const __grid_constant__ Mod<typename C::W> mod) {
    using W = typename C::W;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    W *tile = reinterpret_cast<W *>(smem_raw);
    const int tid = threadIdx.x & (C::P - 1), pl = threadIdx.x >> C::LOGP;
    const size_t poly = (size_t)blockIdx.x * C::PPC + pl;
    const bool active = poly < batch;
    const size_t off = active ? poly * C::N : 0;
    W x[C::R];
    if constexpr (!NATURAL) {
        row_load<C>(x, in + off, tid, active);                  // spectrum order == the coalesced row pattern
    } else if constexpr (C::P <= 32) {
        const int bt = bitrev_n(tid, C::LOGP);
#pragma unroll
        for (int k = 0; k < C::R; ++k) x[k] = active ? ld_stream(in + off + (cbitrev(k, C::LOGR) << C::LOGP) + bt) : (W)0;
    } else {
        // the tile holds the row in natural order: coalesced write, permuted read (the mirror image of the forward store)
        const int bt = bitrev_n(tid, C::LOGP);
        row_load<C>(x, in + off, tid, active);
#pragma unroll
        for (int k = 0; k < C::R; ++k) tile[C::spos(pl * C::N + (k << C::LOGP) + tid)] = x[k];
        tile_sync<C>();
#pragma unroll
        for (int k = 0; k < C::R; ++k) x[k] = tile[C::spos(pl * C::N + ((cbitrev(k, C::LOGR) << C::LOGP) | bt))];
        // this permuted read touches the whole tile, while the first regrouping of dit_all() is ordered by a warp
        // barrier only (exchange_is_warp_local): everybody must be done reading before any warp writes again
        tile_sync<C>();
    }
    dit_all<C, RED, 1, false, C::PREFETCH>(x, tile, pl, tid, inv, TABLE ? post : nullptr, mod);   // canonical input: below one unit
    row_store_scaled<C, TABLE ? 1 : 0, C::POST_GROUP, RED>(x, out + off, tid, active, post, post_uniform, mod);
}

// b_stride = N: one spectrum per row; b_stride = 0: one spectrum shared by the whole batch
template <class C, int RED, int MINB>
__global__ void __launch_bounds__(C::THREADS, MINB)
polymul_spectrum_kernel(const typename C::W *a, const typename C::W *bspec, typename C::W *c, size_t batch, size_t b_stride,
                        const __grid_constant__ PolymulTables<typename C::W> tb,
                        const __grid_constant__ Mod<typename C::W> mod) {
    using W = typename C::W;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    W *tile = reinterpret_cast<W *>(smem_raw);
    const int tid = threadIdx.x & (C::P - 1), pl = threadIdx.x >> C::LOGP;
    const size_t poly = (size_t)blockIdx.x * C::PPC + pl;
    const bool active = poly < batch;
    const size_t off = active ? poly * C::N : 0;
    const W *brow = bspec + (active ? poly * b_stride : 0);
    W x[1][C::R], fa[C::R];
    // b's spectrum is needed one forward transform from now: pull its row towards L2 meanwhile (a spectrum shared
    // by the whole batch is hot anyway)
    for (int line = tid; b_stride != 0 && line < (int)(C::N * sizeof(W) / 128); line += C::P)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(brow) + (size_t)line * 128));
    row_load<C>(x[0], a + off, tid, active);
    forward_all<C, 1, RED, false>(x, tile, pl, tid, tb, mod);
#pragma unroll
    for (int k = 0; k < C::R; ++k) {
        fa[k] = pointwise_product<C, RED>(x[0][k], __ldg(brow + (k << C::LOGP) + tid), mod);   // red 1: < u*q/2^BITS + q, below two units
    }
    dit_all<C, RED, 2, false, C::PREFETCH>(fa, tile, pl, tid, tb.inv, tb.post, mod);
    row_store_scaled<C, 1, C::POST_GROUP, RED>(fa, c + off, tid, active, tb.post, Tw<W>{0, 0}, mod);
}

// The transform-domain kernels over all limbs of a multi-modulus batch in one launch (limb = blockIdx.y, tables out of
// the kernel parameters like polymul_rns_kernel): OP 0 = spectrum_forward (a -> c), 1 = spectrum_inverse (a -> c),
// 2 = polymul_spectrum (a coefficients, b spectra [L][B or 1][N] with row stride b_stride, -> c).
template <class C, int RED, int MINB, int OP>
__global__ void __launch_bounds__(C::THREADS, MINB)
spectrum_rns_kernel(const typename C::W *a, const typename C::W *bspec, typename C::W *c, size_t batch, size_t b_stride,
                    const __grid_constant__ RnsLimbs<typename C::W> limbs) {
    using W = typename C::W;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    W *tile = reinterpret_cast<W *>(smem_raw);
    const PolymulTables<W> &tb = limbs.limb[blockIdx.y].tb;
    const Mod<W> &mod = limbs.limb[blockIdx.y].mod;
    const int tid = threadIdx.x & (C::P - 1), pl = threadIdx.x >> C::LOGP;
    const size_t poly = (size_t)blockIdx.x * C::PPC + pl;
    const bool active = poly < batch;
    const size_t off = (size_t)blockIdx.y * batch * C::N + (active ? poly * C::N : 0);
    if constexpr (OP == 0) {
        W x[1][C::R];
        row_load<C>(x[0], a + off, tid, active);
        forward_all<C, 1, RED, false>(x, tile, pl, tid, tb, mod);
#pragma unroll
        for (int k = 0; k < C::R; ++k) {
            const W v = csub(shoup_mul(x[0][k], (W)1, mod.one_p, mod.nq), mod.q);   // any word -> [0, q)
            if (active) st_stream(c + off + (k << C::LOGP) + tid, v);
        }
    } else if constexpr (OP == 1) {
        const Tw<W> *post = limbs.limb[blockIdx.y].untwist;
        W x[C::R];
        row_load<C>(x, a + off, tid, active);
        dit_all<C, RED, 1, false, C::PREFETCH>(x, tile, pl, tid, tb.inv, post, mod);
        row_store_scaled<C, 1, C::POST_GROUP, RED>(x, c + off, tid, active, post, Tw<W>{0, 0}, mod);
    } else {
        const size_t b_rows = b_stride ? batch : 1;
        const W *brow = bspec + (size_t)blockIdx.y * b_rows * C::N + (active ? poly * b_stride : 0);
        W x[1][C::R], fa[C::R];
        row_load<C>(x[0], a + off, tid, active);
        forward_all<C, 1, RED, false>(x, tile, pl, tid, tb, mod);
#pragma unroll
        for (int k = 0; k < C::R; ++k) fa[k] = pointwise_product<C, RED>(x[0][k], __ldg(brow + (k << C::LOGP) + tid), mod);
        dit_all<C, RED, 2, false, C::PREFETCH>(fa, tile, pl, tid, tb.inv, tb.post, mod);
        row_store_scaled<C, 1, C::POST_GROUP, RED>(fa, c + off, tid, active, tb.post, Tw<W>{0, 0}, mod);
    }
}
// rtl/ntt_pointwise_mult.v:17-42 with the reference's Barrett product, all limbs in one launch (grid.y = limb)
template <typename W>
__global__ void __launch_bounds__(256)
pointwise_rns_kernel(const W *a, const W *b, W *c, size_t count_per_limb, const __grid_constant__ RnsLimbs<W> limbs) {
    const Mod<W> &mod = limbs.limb[blockIdx.y].mod;
    const size_t off = (size_t)blockIdx.y * count_per_limb;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count_per_limb; i += (size_t)gridDim.x * blockDim.x)
        c[off + i] = barrett_mul(a[off + i], b[off + i], mod);
}

// ---------------------------------------------------------------------------------------------
// Small batches: ONE polynomial pair per thread-block CLUSTER (N = 4096)
//
// With fewer rows than SMs the one-CTA-per-row kernel leaves most of the chip idle and a row's latency is
// that of 8 warps working through ~6000 instructions each.  Here the P = N/R threads of a row are spread
// over CS CTAs (CS SMs) of a cluster.  Between passes the coefficients are regrouped through DISTRIBUTED
// shared memory: every thread stores each of its R values straight into the shared memory of the CTA
// whose thread needs it next (mapa + st.shared::cluster), into that thread's private slot [k][thread], so
// the read side is local, contiguous and conflict-free.  Two tile buffers alternate, which makes one
// cluster barrier per exchange sufficient (a buffer is rewritten two exchanges later, after every reader
// has passed the barrier in between).  The twiddles of the next pass are fetched into registers BEFORE the
// barrier, so their L2 latency overlaps the exchange instead of following it.
//   rtl/ntt_coeff_banks.v (ping-pong banks) + rtl/ntt_bank_switch.v -> the two DSMEM tile buffers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Distributed shared memory of a peer CTA may only be touched once that CTA has started executing.  Every CTA
// arrives on the cluster barrier at kernel entry (cluster_entry_arrive) and waits for its peers' arrivals just
// before its first remote store (cluster_entry_wait); the wait hides behind the row load and the first pass.
__device__ __forceinline__ void cluster_entry_arrive() {
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_entry_wait() {
    asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
}
template <typename W> __device__ __forceinline__ void st_cluster(unsigned local_saddr, unsigned rank, W v) {
    unsigned remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_saddr), "r"(rank));
    if constexpr (sizeof(W) == 4) asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote), "r"(v) : "memory");
    else asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(remote), "l"(v) : "memory");
}
// registers (field at LO_FROM) -> the slots of their next owners (field at LO_TO), cluster barrier, local read
template <class C, int CS, int LO_FROM, int LO_TO>
__device__ __forceinline__ void cluster_exchange(typename C::W (&x)[C::R], typename C::W *buf, int gtid) {
    using W = typename C::W;
    constexpr int T = C::P / CS;
    const unsigned base = (unsigned)__cvta_generic_to_shared(buf);
#pragma unroll
    for (int k = 0; k < C::R; ++k) {
        const int E = C::template elem<LO_FROM>(gtid, k);
        const int g2 = ((E >> (LO_TO + C::LOGR)) << LO_TO) | (E & ((1 << LO_TO) - 1));   // next owner (row-wide thread id)
        const int k2 = (E >> LO_TO) & (C::R - 1);                                        // ... and its register
#if defined(TNTT_DEBUG_BOUNDS)
        // -DTNTT_DEBUG_BOUNDS build (make EXTRA=-DTNTT_DEBUG_BOUNDS; compute-sanitizer is closed on the GPU pool): the
        // remote store must land in a CTA of this cluster, inside its N / CS-word buffer, in the slot its reader loads
        assert(g2 >= 0 && g2 / T < CS && k2 * T + (g2 % T) < C::N / CS);
        assert(C::template elem<LO_TO>(g2, k2) == E);
#endif
        st_cluster<W>(base + (unsigned)((k2 * T + (g2 % T)) * sizeof(W)), (unsigned)(g2 / T), x[k]);
    }
    cluster_barrier();
#pragma unroll
    for (int k = 0; k < C::R; ++k) x[k] = buf[k * T + (gtid % T)];
}

// PRELOAD: fetch the next pass's twiddles into registers before the cluster barrier, so that their L2 latency
// overlaps the exchange
template <class C, int CS, int RED, bool PRELOAD, int XI, int PASS = 1>
__device__ __forceinline__ void cluster_forward_rest(typename C::W (&x)[1][C::R], typename C::W *tiles, int gtid,
                                                     const PolymulTables<typename C::W> &tb, const Mod<typename C::W> &mod) {
    if constexpr (PASS < C::NPASS) {
        Tw<typename C::W> tw[PRELOAD ? C::R - 1 : 1];
        if constexpr (PRELOAD) fwd_load_pass_twiddles<C, PASS>(tw, gtid, tb);
        cluster_exchange<C, CS, C::fwd_lo(PASS - 1), C::fwd_lo(PASS)>(x[0], tiles + ((XI + PASS) & 1) * (C::N / CS), gtid);
        fwd_pass<C, PASS, 1, RED, false, false, PRELOAD>(x, gtid, tb, mod, nullptr, tw);
        cluster_forward_rest<C, CS, RED, PRELOAD, XI, PASS + 1>(x, tiles, gtid, tb, mod);
    }
}
template <class C, int CS, int RED, bool PRELOAD, int IN_BND, int XI, int PASS = 1>
__device__ __forceinline__ void cluster_inverse_rest(typename C::W (&x)[C::R], typename C::W *tiles, int gtid,
                                                     const DitTables<typename C::W> &dt, const Mod<typename C::W> &mod) {
    if constexpr (PASS < C::NPASS) {
        Tw<typename C::W> tw[PRELOAD ? C::R - 1 : 1];
        if constexpr (PRELOAD) dit_load_pass_twiddles<C, PASS>(tw, gtid, dt);
        cluster_exchange<C, CS, C::inv_lo(PASS - 1), C::inv_lo(PASS)>(x, tiles + ((XI + PASS) & 1) * (C::N / CS), gtid);
        dit_pass<C, PASS, RED, IN_BND, false, false, PRELOAD>(x, gtid, dt, mod, nullptr, tw);
        cluster_inverse_rest<C, CS, RED, PRELOAD, IN_BND, XI, PASS + 1>(x, tiles, gtid, dt, mod);
    }
}

// launched with a cluster dimension of CS (cudaLaunchKernelEx); grid = rows * CS CTAs of P / CS threads
template <class C, int CS, int RED, int MINB = 1>
__global__ void __launch_bounds__(C::P / CS, MINB)
polymul_cluster_kernel(const typename C::W *a, const typename C::W *b, typename C::W *c,
                       size_t batch, const __grid_constant__ PolymulTables<typename C::W> tb,
                       const __grid_constant__ Mod<typename C::W> mod) {
    using W = typename C::W;
    static_assert(C::PPC == 1 && C::P % CS == 0, "cluster kernel: one row per cluster");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    W *tiles = reinterpret_cast<W *>(smem_raw);                 // two buffers of N / CS words
    const int gtid = (int)cluster_ctarank() * (C::P / CS) + (int)threadIdx.x;
    const size_t row = blockIdx.x / CS;
    const bool active = row < batch;
    const size_t off = active ? row * C::N : 0;
    constexpr int NX = C::NPASS - 1;                            // exchanges per transform
    constexpr bool PRELOAD = true;   // measured: also (slightly) better with two CTAs per SM (3.53 vs 3.32 M/s, 27-bit N = 16384)
#if defined(TNTT_X_EMPTY_KERNEL)
    if (batch) return;   // what-if only: launch overhead calibration
#endif

    cluster_entry_arrive();
    W x[1][C::R], fa[C::R];
    prefetch_post<C>(gtid, tb.post);
    row_load<C>(x[0], a + off, gtid, active);
    row_load<C>(fa, b + off, gtid, active);                      // b's latency overlaps a's transform
    fwd_pass<C, 0, 1, RED>(x, gtid, tb, mod);
    cluster_entry_wait();                                        // every peer CTA is running: its shared memory may be written
    cluster_forward_rest<C, CS, RED, PRELOAD, 0>(x, tiles, gtid, tb, mod);
#pragma unroll
    for (int k = 0; k < C::R; ++k) { const W t = x[0][k]; x[0][k] = fa[k]; fa[k] = t; }
    fwd_pass<C, 0, 1, RED>(x, gtid, tb, mod);
    cluster_forward_rest<C, CS, RED, PRELOAD, NX>(x, tiles, gtid, tb, mod);
#pragma unroll
    for (int k = 0; k < C::R; ++k) {
        fa[k] = pointwise_product<C, RED>(fa[k], x[0][k], mod);
    }
    dit_pass<C, 0, RED, pointwise_out_bound<C, RED>()>(fa, gtid, tb.inv, mod);
    cluster_inverse_rest<C, CS, RED, PRELOAD, pointwise_out_bound<C, RED>(), 2 * NX>(fa, tiles, gtid, tb.inv, mod);
    row_store_scaled<C, 1>(fa, c + off, gtid, active, tb.post, Tw<W>{0, 0}, mod);
    // no trailing barrier: the last remote store into this CTA's shared memory precedes the last exchange's barrier
}

// Transform-domain kernels for rows that live on a cluster (N = 16384, 32768): the halves of polymul_cluster_kernel.
//   MODE 1: spectrum_forward   (in = a, out = c)
//   MODE 2: spectrum_inverse   (in = a spectrum, out = c; post = psi^-i N^-1 table)
//   MODE 3: polymul_spectrum   (a coefficients, b spectrum with row stride b_stride, out = c; post = tb.post)
// The spectrum order is the same as in the one-CTA kernels: word k*P + t = bit-reversed-order element t*R + k, with t
// the row-wide thread index (cluster rank * threads + threadIdx).
template <class C, int CS, int RED, int MINB, int MODE>
__global__ void __launch_bounds__(C::P / CS, MINB)
spectrum_cluster_kernel(const typename C::W *a, const typename C::W *b, typename C::W *c,
                        size_t batch, size_t b_stride, const __grid_constant__ PolymulTables<typename C::W> tb,
                        const Tw<typename C::W> *__restrict__ post, const __grid_constant__ Mod<typename C::W> mod) {
    using W = typename C::W;
    static_assert(C::PPC == 1 && C::P % CS == 0 && MODE >= 1 && MODE <= 3, "cluster kernel: one row per cluster");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    W *tiles = reinterpret_cast<W *>(smem_raw);
    const int gtid = (int)cluster_ctarank() * (C::P / CS) + (int)threadIdx.x;
    const size_t row = blockIdx.x / CS;
    const bool active = row < batch;
    const size_t off = active ? row * C::N : 0;
    constexpr int NX = C::NPASS - 1;
    cluster_entry_arrive();                                      // see polymul_cluster_kernel
    W fa[C::R];
    if constexpr (MODE == 2) {
        row_load<C>(fa, a + off, gtid, active);
        dit_pass<C, 0, RED, 1>(fa, gtid, tb.inv, mod);
        cluster_entry_wait();
        cluster_inverse_rest<C, CS, RED, true, 1, 0>(fa, tiles, gtid, tb.inv, mod);
        row_store_scaled<C, 1>(fa, c + off, gtid, active, post, Tw<W>{0, 0}, mod);
    } else {
        W x[1][C::R];
        const W *brow = b + (active ? row * b_stride : 0);
        if constexpr (MODE == 3) {
            for (int line = gtid; b_stride != 0 && line < (int)(C::N * sizeof(W) / 128); line += C::P)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(brow) + (size_t)line * 128));
        }
        row_load<C>(x[0], a + off, gtid, active);
        fwd_pass<C, 0, 1, RED>(x, gtid, tb, mod);
        cluster_entry_wait();
        cluster_forward_rest<C, CS, RED, true, 0>(x, tiles, gtid, tb, mod);
        if constexpr (MODE == 1) {
#pragma unroll
            for (int k = 0; k < C::R; ++k) {
                const W v = csub(shoup_mul(x[0][k], (W)1, mod.one_p, mod.nq), mod.q);
                if (active) st_stream(c + off + (k << C::LOGP) + gtid, v);
            }
        } else {
#pragma unroll
            for (int k = 0; k < C::R; ++k) {
                fa[k] = pointwise_product<C, RED>(x[0][k], __ldg(brow + (k << C::LOGP) + gtid), mod);
            }
            dit_pass<C, 0, RED, 2>(fa, gtid, tb.inv, mod);
            cluster_inverse_rest<C, CS, RED, true, 2, NX>(fa, tiles, gtid, tb.inv, mod);
            row_store_scaled<C, 1>(fa, c + off, gtid, active, post, Tw<W>{0, 0}, mod);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// standalone natural-order transform (cg_ntt / cg_intt semantics, new_reference/cg_ntt.py:29-75):
//   out = post * DFT_root(pre * in), all in natural order.  The bit-reversal of cg_ntt.py:39 /
//   rtl/ntt_coeff_banks.v:112 happens on the way into the shared tile.
// ---------------------------------------------------------------------------------------------
template <class C, int RED, int MINB>
__global__ void __launch_bounds__(C::THREADS, MINB)
transform_kernel(const typename C::W *in, typename C::W *out, size_t batch,
                 const __grid_constant__ TransformTables<typename C::W> tb,
                 const __grid_constant__ Mod<typename C::W> mod) {
    using W = typename C::W;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    W *tile = reinterpret_cast<W *>(smem_raw);
    const int tid = threadIdx.x & (C::P - 1), pl = threadIdx.x >> C::LOGP;
    const size_t poly = (size_t)blockIdx.x * C::PPC + pl;
    const bool active = poly < batch;
    const size_t off = active ? poly * C::N : 0;

    W x[C::R];
    row_load<C>(x, in + off, tid, active);
#pragma unroll
    for (int k = 0; k < C::R; ++k) {
        const int e = (k << C::LOGP) + tid;
        if (tb.pre) x[k] = shoup_mul(x[k], ld_tw(&tb.pre[e]), mod.nq);
        else if (tb.reduce_input) x[k] = shoup_mul(x[k], (W)1, mod.one_p, mod.nq);
        tile[C::spos(pl * C::N + bitrev_n(e, C::LOGN))] = x[k];
    }
    tile_sync<C>();
    tile_read<C, 0>(x, tile, pl, tid);
    dit_all<C, RED, 2, false, true>(x, tile, pl, tid, tb.dit, tb.post, mod);   // short kernel: latency-bound without it (measured 2x)
    row_store_scaled<C>(x, out + off, tid, active, tb.post, tb.post_uniform, mod);
}
#endif  // __CUDACC__

}  // namespace tntt
