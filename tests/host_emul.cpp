// CPU emulation of the CUDA kernels' thread/tile choreography.  TEST FIXTURE ONLY.
//
// Includes the very same __host__ __device__ building blocks the kernels are made of
// (tiny-ntt_b200/csrc/kernels.cuh: index maps, swizzle, passes, arithmetic) and replays the
// kernel bodies with "for every thread" loops between the __syncthreads() points.  This lets
// the CPU-only test-suite (-m "not gpu") check index maps, tables and lazy-range arithmetic
// bit-for-bit against the oracle without a GPU.  It is not part of the product library and is
// never used as a fallback; it also records the shared-memory slots touched per warp
// instruction so tests/test_layout.py can count bank conflicts.
#define TNTT_AUDIT_RANGES 1
#include <cstdio>
#include <cstring>
#include <vector>

#include "../tiny-ntt_b200/csrc/kernels.cuh"
#include "../tiny-ntt_b200/csrc/tables.h"

using namespace tntt;
namespace tntt { long long g_range_violations = 0; }

namespace {

template <class C, int NA, int RED> struct Emu {
    using W = typename C::W;
    static constexpr int T = C::THREADS;
    std::vector<W> tile;
    std::vector<W> x;   // [T][NA][R]
    std::vector<W> fa;  // [T][R]
    PolymulTables<W> tb;
    Mod<W> mod;
    long long overflow_checks = 0;

    W (&X(int t))[NA][C::R] { return *reinterpret_cast<W(*)[NA][C::R]>(&x[(size_t)t * NA * C::R]); }
    W (&F(int t))[C::R] { return *reinterpret_cast<W(*)[C::R]>(&fa[(size_t)t * C::R]); }

    template <int PASS> void forward_from() {
        if constexpr (PASS < C::NPASS) {
            if constexpr (PASS > 0) {
                for (int a = 0; a < NA; ++a) {
                    W *tl = tile.data() + (size_t)a * C::TILE;
                    for (int t = 0; t < T; ++t)
                        tile_write<C, C::fwd_lo(PASS - 1)>(X(t)[a], tl, t >> C::LOGP, t & (C::P - 1));
                    for (int t = 0; t < T; ++t) {
                        tile_read<C, C::fwd_lo(PASS)>(X(t)[a], tl, t >> C::LOGP, t & (C::P - 1));
                    }
                }
            }
            for (int t = 0; t < T; ++t) fwd_pass<C, PASS, NA, RED>(X(t), t & (C::P - 1), tb, mod);
            forward_from<PASS + 1>();
        }
    }
    template <int IN_BND, int PASS> void dit_from(const DitTables<W> &pyr) {
        if constexpr (PASS < C::NPASS) {
            if constexpr (PASS > 0) {
                for (int t = 0; t < T; ++t)
                    tile_write<C, C::inv_lo(PASS - 1)>(F(t), tile.data(), t >> C::LOGP, t & (C::P - 1));
                for (int t = 0; t < T; ++t) {
                    tile_read<C, C::inv_lo(PASS)>(F(t), tile.data(), t >> C::LOGP, t & (C::P - 1));
                }
            }
            for (int t = 0; t < T; ++t) dit_pass<C, PASS, RED, IN_BND>(F(t), t & (C::P - 1), pyr, mod);
            dit_from<IN_BND, PASS + 1>(pyr);
        }
    }

    void polymul(const W *a, const W *b, W *c, size_t batch) {
        tile.assign((size_t)NA * C::TILE, 0);
        x.assign((size_t)T * NA * C::R, 0);
        fa.assign((size_t)T * C::R, 0);
        const size_t ctas = (batch + C::PPC - 1) / C::PPC;
        for (size_t cta = 0; cta < ctas; ++cta) {
            auto off = [&](int t, bool &active) {
                const size_t poly = cta * C::PPC + (t >> C::LOGP);
                active = poly < batch;
                return active ? poly * C::N : 0;
            };
            bool act;
            if (NA == 1) {
                for (int t = 0; t < T; ++t) { size_t o = off(t, act); row_load<C>(X(t)[0], a + o, t & (C::P - 1), act); }
                forward_from<0>();
                for (int t = 0; t < T; ++t) for (int k = 0; k < C::R; ++k) F(t)[k] = X(t)[0][k];
                for (int t = 0; t < T; ++t) { size_t o = off(t, act); row_load<C>(X(t)[0], b + o, t & (C::P - 1), act); }
                forward_from<0>();
                for (int t = 0; t < T; ++t)
                    for (int k = 0; k < C::R; ++k) {
                        F(t)[k] = pointwise_product<C, RED>(F(t)[k], X(t)[0][k], mod);
                    }
            } else {
                for (int t = 0; t < T; ++t) {
                    size_t o = off(t, act);
                    row_load<C>(X(t)[0], a + o, t & (C::P - 1), act);
                    row_load<C>(X(t)[NA - 1], b + o, t & (C::P - 1), act);
                }
                forward_from<0>();
                for (int t = 0; t < T; ++t)
                    for (int k = 0; k < C::R; ++k) {
                        F(t)[k] = pointwise_product<C, RED>(X(t)[0][k], X(t)[NA - 1][k], mod);
                    }
            }
            dit_from<pointwise_out_bound<C, RED>(), 0>(tb.inv);
            for (int t = 0; t < T; ++t) {
                size_t o = off(t, act);
                row_store_scaled<C, 1, C::POST_GROUP, RED>(F(t), c + o, t & (C::P - 1), act, tb.post, Tw<W>{0, 0}, mod);
            }
        }
    }

    // transform-domain kernels (spectrum_forward_kernel / spectrum_inverse_kernel / polymul_spectrum_kernel)
    void spectrum_forward(const W *in, W *out, size_t batch, bool natural = false) {
        tile.assign((size_t)C::TILE, 0);
        x.assign((size_t)T * NA * C::R, 0);
        const size_t ctas = (batch + C::PPC - 1) / C::PPC;
        for (size_t cta = 0; cta < ctas; ++cta) {
            for (int t = 0; t < T; ++t) {
                const size_t poly = cta * C::PPC + (t >> C::LOGP);
                row_load<C>(X(t)[0], in + (poly < batch ? poly * C::N : 0), t & (C::P - 1), poly < batch);
            }
            forward_from<0>();
            for (int t = 0; t < T; ++t) {
                const size_t poly = cta * C::PPC + (t >> C::LOGP);
                if (poly >= batch) continue;
                for (int k = 0; k < C::R; ++k) {
                    const int tid = t & (C::P - 1);
                    const size_t idx = natural ? (size_t)((cbitrev(k, C::LOGR) << C::LOGP) | bitrev_n(tid, C::LOGP)) : (size_t)((k << C::LOGP) + tid);
                    out[poly * C::N + idx] = csub(shoup_mul(X(t)[0][k], (W)1, mod.one_p, mod.nq), mod.q);
                }
            }
        }
    }
    void spectrum_inverse(const W *in, W *out, size_t batch, const Tw<W> *post, bool natural = false, Tw<W> uniform = Tw<W>{0, 0}) {
        tile.assign((size_t)C::TILE, 0);
        fa.assign((size_t)T * C::R, 0);
        const size_t ctas = (batch + C::PPC - 1) / C::PPC;
        for (size_t cta = 0; cta < ctas; ++cta) {
            for (int t = 0; t < T; ++t) {
                const size_t poly = cta * C::PPC + (t >> C::LOGP);
                if (!natural) row_load<C>(F(t), in + (poly < batch ? poly * C::N : 0), t & (C::P - 1), poly < batch);
                else
                    for (int k = 0; k < C::R; ++k)
                        F(t)[k] = poly < batch ? in[poly * C::N + ((cbitrev(k, C::LOGR) << C::LOGP) | bitrev_n(t & (C::P - 1), C::LOGP))] : (W)0;
            }
            dit_from<1, 0>(tb.inv);
            for (int t = 0; t < T; ++t) {
                const size_t poly = cta * C::PPC + (t >> C::LOGP);
                if (post) row_store_scaled<C, 1, C::POST_GROUP, RED>(F(t), out + (poly < batch ? poly * C::N : 0), t & (C::P - 1), poly < batch, post, Tw<W>{0, 0}, mod);
                else row_store_scaled<C, 0>(F(t), out + (poly < batch ? poly * C::N : 0), t & (C::P - 1), poly < batch, nullptr, uniform, mod);
            }
        }
    }
    void polymul_spectrum(const W *a, const W *bspec, W *c, size_t batch, size_t b_stride) {
        tile.assign((size_t)C::TILE, 0);
        x.assign((size_t)T * NA * C::R, 0);
        fa.assign((size_t)T * C::R, 0);
        const size_t ctas = (batch + C::PPC - 1) / C::PPC;
        for (size_t cta = 0; cta < ctas; ++cta) {
            for (int t = 0; t < T; ++t) {
                const size_t poly = cta * C::PPC + (t >> C::LOGP);
                row_load<C>(X(t)[0], a + (poly < batch ? poly * C::N : 0), t & (C::P - 1), poly < batch);
            }
            forward_from<0>();
            for (int t = 0; t < T; ++t) {
                const size_t poly = cta * C::PPC + (t >> C::LOGP);
                const W *brow = bspec + (poly < batch ? poly * b_stride : 0);
                for (int k = 0; k < C::R; ++k) {
                    F(t)[k] = pointwise_product<C, RED>(X(t)[0][k], brow[(k << C::LOGP) + (t & (C::P - 1))], mod);
                }
            }
            dit_from<2, 0>(tb.inv);
            for (int t = 0; t < T; ++t) {
                const size_t poly = cta * C::PPC + (t >> C::LOGP);
                row_store_scaled<C, 1, C::POST_GROUP, RED>(F(t), c + (poly < batch ? poly * C::N : 0), t & (C::P - 1), poly < batch, tb.post, Tw<W>{0, 0}, mod);
            }
        }
    }

    // standalone transform kernel body (transform_kernel in kernels.cuh)
    void transform(const W *in, W *out, size_t batch, const TransformTables<W> &tt) {
        tile.assign((size_t)C::TILE, 0);
        fa.assign((size_t)T * C::R, 0);
        const size_t ctas = (batch + C::PPC - 1) / C::PPC;
        for (size_t cta = 0; cta < ctas; ++cta) {
            for (int t = 0; t < T; ++t) {
                const int tid = t & (C::P - 1), pl = t >> C::LOGP;
                const size_t poly = cta * C::PPC + pl;
                const bool active = poly < batch;
                const size_t o = active ? poly * C::N : 0;
                row_load<C>(F(t), in + o, tid, active);
                for (int k = 0; k < C::R; ++k) {
                    const int e = (k << C::LOGP) + tid;
                    if (tt.pre) F(t)[k] = shoup_mul(F(t)[k], ld_tw(&tt.pre[e]), mod.nq);
                    else if (tt.reduce_input) F(t)[k] = shoup_mul(F(t)[k], (W)1, mod.one_p, mod.nq);
                    tile[C::spos(pl * C::N + bitrev_n(e, C::LOGN))] = F(t)[k];
                }
            }
            for (int t = 0; t < T; ++t) tile_read<C, 0>(F(t), tile.data(), t >> C::LOGP, t & (C::P - 1));
            dit_from<2, 0>(tt.dit);
            for (int t = 0; t < T; ++t) {
                const size_t poly = cta * C::PPC + (t >> C::LOGP);
                const bool active = poly < batch;
                row_store_scaled<C>(F(t), out + (active ? poly * C::N : 0), t & (C::P - 1), active, tt.post,
                                    tt.post_uniform, mod);
            }
        }
    }
};

template <class C, int NA, int RED>
int run_polymul(const void *a, const void *b, void *c, size_t batch, uint64_t q, uint64_t psi) {
    using W = typename C::W;
    constexpr int BITS = WordTraits<W>::BITS;
    if (RED == 3 ? q >= (1ull << 60) : (RED ? !host::lazy_pass_ok<W>(q, C::LOGR) : !host::lazy_full_ok<W>(q, C::LOGN))) return -2;
    if (RED == 2 && q != kSolinasQ) return -2;
    const uint64_t omega = host::mulmod(psi, psi, q);
    auto fwd = host::fwd_pyramid<W>(psi, C::N, q);
    auto last = host::fwd_last_table<W>(fwd, C::LOGN, C::LOGR);
    auto inv = host::dit_pyramid<W>(host::modinv(omega, q), C::N, q);
    // the Montgomery pointwise product (red 0/1) leaves a factor 2^-BITS for the store table to undo; the Solinas one does not
    const uint64_t scale = RED >= 2 ? host::modinv(C::N % q, q)
                                    : host::mulmod(host::modinv(C::N % q, q), (uint64_t)((((host::u128)1) << BITS) % q), q);
    auto post = host::scaled_powers<W>(host::modinv(psi, q), scale, C::N, q);
    Emu<C, NA, RED> e;
    e.tb.fwd_pyr = fwd.data();
    e.tb.fwd_last = last.data();
    e.tb.post = post.data();
    e.tb.inv.pyr = inv.data();
    for (int i = 0; i < MAX_R && i < C::N; ++i) { e.tb.fwd_head[i] = fwd[i]; e.tb.inv.head[i] = inv[i]; }
    e.mod = host::make_mod<W>(q, C::LOGN);
    e.polymul((const W *)a, (const W *)b, (W *)c, batch);
    return 0;
}

// mode 0: spec = forward(a), out = inverse(spec); 1: out = polymul_spectrum(a, forward(b)), one spectrum per row;
// 2: the same with the spectrum of b's row 0 shared by the batch; 3: out = forward(a) (the raw spectrum)
template <class C, int RED>
int run_spectrum(const void *a, const void *b, void *out, size_t batch, uint64_t q, uint64_t psi, int mode) {
    using W = typename C::W;
    constexpr int BITS = WordTraits<W>::BITS;
    if (RED == 3 ? q >= (1ull << 60) : (RED ? !host::lazy_pass_ok<W>(q, C::LOGR) : !host::lazy_full_ok<W>(q, C::LOGN))) return -2;
    const uint64_t omega = host::mulmod(psi, psi, q), n_inv = host::modinv(C::N % q, q);
    auto fwd = host::fwd_pyramid<W>(psi, C::N, q);
    auto last = host::fwd_last_table<W>(fwd, C::LOGN, C::LOGR);
    auto inv = host::dit_pyramid<W>(host::modinv(omega, q), C::N, q);
    auto post_mont = host::scaled_powers<W>(host::modinv(psi, q), host::mulmod(n_inv, (uint64_t)((((host::u128)1) << BITS) % q), q), C::N, q);
    auto post_plain = host::scaled_powers<W>(host::modinv(psi, q), n_inv, C::N, q);
    if (RED == 2 && q != kSolinasQ) return -2;
    Emu<C, 1, RED> e;
    e.tb.fwd_pyr = fwd.data();
    e.tb.fwd_last = last.data();
    e.tb.post = RED == 2 ? post_plain.data() : post_mont.data();   // the Solinas pointwise product leaves no 2^-BITS
    e.tb.inv.pyr = inv.data();
    for (int i = 0; i < MAX_R && i < C::N; ++i) { e.tb.fwd_head[i] = fwd[i]; e.tb.inv.head[i] = inv[i]; }
    e.mod = host::make_mod<W>(q, C::LOGN);
    std::vector<W> spec((size_t)batch * C::N);
    if (mode >= 4) {   // natural-order transforms: 4 = ntt(twist(a)), 5 = cg_ntt(a), 6 = twisted inverse, 7 = cg_intt
        auto cyc = host::fwd_pyramid_cyclic<W>(omega, C::N, q);
        auto cyc_last = host::fwd_last_table<W>(cyc, C::LOGN, C::LOGR);
        if (mode == 5) {
            e.tb.fwd_pyr = cyc.data();
            e.tb.fwd_last = cyc_last.data();
            for (int i = 0; i < MAX_R && i < C::N; ++i) e.tb.fwd_head[i] = cyc[i];
        }
        if (mode <= 5) e.spectrum_forward((const W *)a, (W *)out, batch, true);
        else e.spectrum_inverse((const W *)a, (W *)out, batch, mode == 6 ? post_plain.data() : nullptr, true, host::make_tw<W>(n_inv, q));
        return 0;
    }
    if (mode == 0 || mode == 3) {
        e.spectrum_forward((const W *)a, mode == 3 ? (W *)out : spec.data(), batch);
        if (mode == 0) e.spectrum_inverse(spec.data(), (W *)out, batch, post_plain.data());
    } else {
        e.spectrum_forward((const W *)b, spec.data(), mode == 2 ? 1 : batch);
        e.polymul_spectrum((const W *)a, spec.data(), (W *)out, batch, mode == 2 ? 0 : C::N);
    }
    return 0;
}

// mode 0: cg_ntt (root = omega); 1: cg_intt; 2: twisted forward; 3: twisted inverse
template <class C, bool RED>
int run_transform(const void *in, void *out, size_t batch, uint64_t q, uint64_t root, int mode, int reduce_input) {
    using W = typename C::W;
    const bool inverse = mode & 1, twist = mode & 2;
    const uint64_t omega = twist ? host::mulmod(root, root, q) : root;
    auto pyr = host::dit_pyramid<W>(inverse ? host::modinv(omega, q) : omega, C::N, q);
    std::vector<Tw<W>> pre, post;
    const uint64_t n_inv = host::modinv(C::N % q, q);
    if (twist && !inverse) pre = host::scaled_powers<W>(root, 1, C::N, q);
    if (twist && inverse) post = host::scaled_powers<W>(host::modinv(root, q), n_inv, C::N, q);
    TransformTables<W> tt;
    tt.dit.pyr = pyr.data();
    for (int i = 0; i < MAX_R && i < (int)pyr.size(); ++i) tt.dit.head[i] = pyr[i];
    tt.pre = pre.empty() ? nullptr : pre.data();
    tt.post = post.empty() ? nullptr : post.data();
    tt.post_uniform = host::make_tw<W>(inverse ? n_inv : 1, q);
    tt.reduce_input = reduce_input;
    Emu<C, 1, RED> e;
    e.mod = host::make_mod<W>(q, C::LOGN);
    e.transform((const W *)in, (W *)out, batch, tt);
    return 0;
}

}  // namespace

#define POLY_CASE(WB, WT, LN, LR, PPC, NA_, RED_) POLY_CASE_P(WB, WT, LN, LR, PPC, NA_, RED_, 0)
#define POLY_CASE_P(WB, WT, LN, LR, PPC, NA_, RED_, PAD_)                                              \
    if (word_bytes == WB && logn == LN && logr == LR && ppc == PPC && na == NA_ && red == RED_ && pad == PAD_) \
        return run_polymul<Cfg<WT, LN, LR, PPC, PAD_>, NA_, RED_>(a, b, c, batch, q, psi);
#define XFORM_CASE(WB, WT, LN, LR, PPC, RED_)                                                          \
    if (word_bytes == WB && logn == LN && logr == LR && ppc == PPC && red == RED_)                    \
        return run_transform<Cfg<WT, LN, LR, PPC>, (RED_ != 0)>(in, out, batch, q, root, mode, reduce_input);

#define SPEC_CASE(WB, WT, LN, LR, PPC, RED_)                                                           \
    if (word_bytes == WB && logn == LN && logr == LR && ppc == PPC && red == RED_)                    \
        return run_spectrum<Cfg<WT, LN, LR, PPC>, RED_>(a, b, out, batch, q, psi, mode);

extern "C" {

// the shapes of tiny-ntt_b200/csrc/spectrum.cu
int emu_spectrum(int word_bytes, int logn, int logr, int ppc, int red, const void *a, const void *b, void *out, size_t batch,
                 uint64_t q, uint64_t psi, int mode) {
    SPEC_CASE(4, uint32_t, 8, 4, 16, 0)
    SPEC_CASE(4, uint32_t, 10, 5, 8, 0)
    SPEC_CASE(4, uint32_t, 12, 4, 1, 0)
    SPEC_CASE(8, uint64_t, 8, 4, 16, 0)
    SPEC_CASE(8, uint64_t, 8, 4, 16, 1)
    SPEC_CASE(8, uint64_t, 10, 4, 4, 0)
    SPEC_CASE(8, uint64_t, 10, 4, 4, 1)
    SPEC_CASE(8, uint64_t, 12, 4, 1, 0)
    SPEC_CASE(8, uint64_t, 12, 4, 1, 1)
    SPEC_CASE(8, uint64_t, 12, 4, 1, 2)
    SPEC_CASE(4, uint32_t, 9, 5, 16, 0)
    SPEC_CASE(4, uint32_t, 11, 4, 2, 0)
    SPEC_CASE(4, uint32_t, 13, 5, 1, 0)
    SPEC_CASE(8, uint64_t, 9, 4, 8, 0)
    SPEC_CASE(8, uint64_t, 9, 4, 8, 1)
    SPEC_CASE(8, uint64_t, 11, 4, 2, 0)
    SPEC_CASE(8, uint64_t, 11, 4, 2, 1)
    SPEC_CASE(8, uint64_t, 13, 4, 1, 0)
    SPEC_CASE(8, uint64_t, 13, 4, 1, 1)
    return -1;
}

// red: 0 / 1 / 2 (Solinas, q = 2^60 - 2^14 + 1 only); pad: 1 = padded tile
int emu_polymul_ex(int word_bytes, int logn, int logr, int ppc, int na, int red, int pad, const void *a, const void *b, void *c,
                   size_t batch, uint64_t q, uint64_t psi) {
    POLY_CASE_P(4, uint32_t, 8, 4, 16, 2, 0, 1)
    POLY_CASE_P(4, uint32_t, 10, 5, 8, 2, 0, 1)
    POLY_CASE_P(4, uint32_t, 12, 4, 1, 2, 0, 1)
    POLY_CASE_P(8, uint64_t, 12, 4, 1, 1, 3, 0)
    POLY_CASE_P(8, uint64_t, 12, 4, 1, 2, 3, 1)
    POLY_CASE_P(8, uint64_t, 8, 4, 16, 1, 3, 0)
    POLY_CASE_P(8, uint64_t, 12, 4, 1, 1, 1, 1)
    POLY_CASE_P(8, uint64_t, 12, 4, 1, 2, 1, 1)
    POLY_CASE_P(8, uint64_t, 12, 4, 1, 1, 2, 0)
    POLY_CASE_P(8, uint64_t, 12, 4, 1, 2, 2, 0)
    POLY_CASE_P(8, uint64_t, 12, 4, 1, 1, 2, 1)
    POLY_CASE_P(8, uint64_t, 12, 4, 1, 2, 2, 1)
    POLY_CASE_P(8, uint64_t, 8, 4, 16, 1, 2, 0)     // small Solinas shapes: quick exhaustive-ish tests
    POLY_CASE_P(8, uint64_t, 8, 4, 16, 1, 1, 1)
    POLY_CASE_P(8, uint64_t, 10, 4, 4, 1, 2, 1)
    POLY_CASE(4, uint32_t, 2, 1, 2, 1, 0)
    POLY_CASE(4, uint32_t, 4, 2, 4, 1, 0)
    POLY_CASE(4, uint32_t, 5, 2, 2, 2, 0)
    POLY_CASE(4, uint32_t, 8, 4, 16, 1, 0)
    POLY_CASE(4, uint32_t, 8, 4, 16, 2, 0)
    POLY_CASE(4, uint32_t, 8, 3, 8, 1, 0)
    POLY_CASE(4, uint32_t, 10, 5, 8, 1, 0)
    POLY_CASE(4, uint32_t, 10, 5, 8, 2, 0)
    POLY_CASE(4, uint32_t, 10, 4, 4, 1, 0)
    POLY_CASE(4, uint32_t, 10, 4, 4, 2, 0)
    POLY_CASE(4, uint32_t, 12, 4, 1, 1, 0)
    POLY_CASE(4, uint32_t, 12, 4, 1, 2, 0)
    POLY_CASE(4, uint32_t, 12, 5, 2, 1, 0)
    POLY_CASE(4, uint32_t, 12, 3, 1, 1, 0)
    POLY_CASE(8, uint64_t, 8, 4, 16, 1, 0)
    POLY_CASE(8, uint64_t, 8, 4, 16, 1, 1)
    POLY_CASE(8, uint64_t, 10, 4, 4, 1, 0)
    POLY_CASE(8, uint64_t, 10, 4, 4, 1, 1)
    POLY_CASE(8, uint64_t, 12, 4, 1, 1, 0)
    POLY_CASE(8, uint64_t, 12, 4, 1, 1, 1)
    POLY_CASE(8, uint64_t, 12, 4, 1, 2, 1)
    POLY_CASE(8, uint64_t, 12, 3, 1, 1, 1)
    POLY_CASE(8, uint64_t, 12, 3, 1, 2, 1)
    // N = 512, 2048, 8192
    POLY_CASE(4, uint32_t, 9, 5, 16, 2, 0)
    POLY_CASE(4, uint32_t, 11, 4, 2, 2, 0)
    POLY_CASE(4, uint32_t, 13, 5, 1, 2, 0)
    POLY_CASE(8, uint64_t, 9, 4, 8, 1, 0)
    POLY_CASE(8, uint64_t, 9, 4, 8, 1, 1)
    POLY_CASE(8, uint64_t, 11, 4, 2, 1, 0)
    POLY_CASE(8, uint64_t, 11, 4, 2, 1, 1)
    POLY_CASE(8, uint64_t, 11, 4, 2, 2, 1)
    POLY_CASE(8, uint64_t, 13, 4, 1, 1, 0)
    POLY_CASE(8, uint64_t, 13, 4, 1, 1, 1)
    return -1;
}

int emu_polymul(int word_bytes, int logn, int logr, int ppc, int na, int red, const void *a, const void *b, void *c,
                size_t batch, uint64_t q, uint64_t psi) {
    return emu_polymul_ex(word_bytes, logn, logr, ppc, na, red, 0, a, b, c, batch, q, psi);
}

int emu_transform(int word_bytes, int logn, int logr, int ppc, int red, const void *in, void *out, size_t batch,
                  uint64_t q, uint64_t root, int mode, int reduce_input) {
    XFORM_CASE(4, uint32_t, 4, 2, 4, 0)
    XFORM_CASE(4, uint32_t, 8, 4, 16, 0)
    XFORM_CASE(4, uint32_t, 10, 5, 8, 0)
    XFORM_CASE(4, uint32_t, 10, 4, 4, 0)
    XFORM_CASE(4, uint32_t, 12, 4, 1, 0)
    XFORM_CASE(8, uint64_t, 8, 4, 16, 1)
    XFORM_CASE(8, uint64_t, 12, 4, 1, 0)
    XFORM_CASE(8, uint64_t, 12, 4, 1, 1)
    XFORM_CASE(8, uint64_t, 12, 3, 1, 1)
    return -1;
}

// Distributed-shared-memory addressing of cluster_exchange() (kernels.cuh), replayed for every thread of a row: each
// (destination CTA, slot) must lie inside the cluster and its N / CS-word buffer, be written exactly once, and hold the
// coefficient its reader -- thread g2 of the row, register k2, slot k2 * T + g2 % T of CTA g2 / T -- owns in the next
// layout.  Returns the number of violations.
int emu_cluster_exchange_violations(int logn, int logr, int cs, int lo_from, int lo_to) {
    const int N = 1 << logn, R = 1 << logr, P = N / R, T = P / cs;
    if (T * cs != P) return -1;
    auto elem = [&](int lo, int tid, int k) { return ((tid >> lo) << (lo + logr)) | (k << lo) | (tid & ((1 << lo) - 1)); };
    std::vector<int> buf((size_t)cs * (N / cs), -1);
    int bad = 0;
    for (int gtid = 0; gtid < P; ++gtid)
        for (int k = 0; k < R; ++k) {
            const int E = elem(lo_from, gtid, k);
            const int g2 = ((E >> (lo_to + logr)) << lo_to) | (E & ((1 << lo_to) - 1));
            const int k2 = (E >> lo_to) & (R - 1);
            const int cta = g2 / T, slot = k2 * T + (g2 % T);
            if (cta < 0 || cta >= cs || slot < 0 || slot >= N / cs) { ++bad; continue; }
            if (buf[(size_t)cta * (N / cs) + slot] != -1) ++bad;          // two writers
            buf[(size_t)cta * (N / cs) + slot] = E;
        }
    for (int gtid = 0; gtid < P; ++gtid)                                   // the read side: buf[k * T + gtid % T] of the own CTA
        for (int k = 0; k < R; ++k)
            if (buf[(size_t)(gtid / T) * (N / cs) + k * T + (gtid % T)] != elem(lo_to, gtid, k)) ++bad;
    return bad;
}

// shared-memory slot of (poly-in-cta, register k, thread tid) for a register field at bit `lo`
int emu_slot_ex(int word_bytes, int logn, int logr, int lo, int pl, int tid, int k, int pad) {
    const int n = 1 << logn;
    const int e = ((tid >> lo) << (lo + logr)) | (k << lo) | (tid & ((1 << lo) - 1));
    const int E = pl * n + e;
    if (pad) return E + (E >> logr);   // Cfg::spos, PAD = 1
    const int mask = (1 << (word_bytes == 4 ? 5 : 4)) - 1;
    return E ^ ((E >> logr) & mask);
}
int emu_slot(int word_bytes, int logn, int logr, int lo, int pl, int tid, int k) { return emu_slot_ex(word_bytes, logn, logr, lo, pl, tid, k, 0); }
uint64_t emu_solinas_reduce(uint64_t x) { return solinas_reduce(x); }
uint64_t emu_solinas_mul(uint64_t u, uint64_t v) { return solinas_mul(u, v); }

// arithmetic probes for tests/test_modarith.py
uint64_t emu_shoup64(uint64_t x, uint64_t w, uint64_t q) { auto t = host::make_tw<uint64_t>(w, q); return shoup_mul(x, t.w, t.wp, (uint64_t)(0 - q)); }
uint32_t emu_shoup32(uint32_t x, uint32_t w, uint32_t q) { auto t = host::make_tw<uint32_t>(w, q); return shoup_mul(x, t.w, t.wp, (uint32_t)(0u - q)); }
uint64_t emu_shoup_lazy64(uint64_t x, uint64_t w, uint64_t q) { auto t = host::make_tw<uint64_t>(w, q); return shoup_lazy(x, t.w, t.wp, host::make_mod<uint64_t>(q)); }
uint64_t emu_mont64(uint64_t x, uint64_t y, uint64_t q) { return mont_mul(x, y, host::make_mod<uint64_t>(q)); }
uint32_t emu_mont32(uint32_t x, uint32_t y, uint32_t q) { return mont_mul(x, y, host::make_mod<uint32_t>(q)); }
uint64_t emu_barrett64(uint64_t x, uint64_t y, uint64_t q) { return barrett_mul(x, y, host::make_mod<uint64_t>(q)); }
uint32_t emu_barrett32(uint32_t x, uint32_t y, uint32_t q) { return barrett_mul(x, y, host::make_mod<uint32_t>(q)); }
uint64_t emu_csub_top64(uint64_t x, uint64_t q) { return csub_top(x, host::make_mod<uint64_t>(q).top_sub); }
long long emu_range_violations(void) { return tntt::g_range_violations; }
int emu_is_prime(uint64_t n) { return host::is_prime(n); }
// the per-register bound tracker of the Solinas kernels' first inverse pass (modarith.cuh): bounds entering `stage`, the
// uniform bound every later stage starts from, the decisions of one butterfly, and the last stage that takes the shortcut
void emu_dit2_pass0_bounds(int g, int b0, int logr, int stage, int *out) {
    const RegBounds r = dit2_pass0_bounds(g, b0, logr, stage);
    for (int k = 0; k < (1 << logr); ++k) out[k] = r.b[k];
}
int emu_dit2_bound_at(int g, int b0, int logr, int stage) { return dit2_bound_at(g, b0, logr, stage); }
int emu_dit2_step(int trivial, int g, int bx, int by, int *out4) {
    const Dit2Step s = dit2_step(trivial != 0, g, bx, by);
    out4[0] = s.bx; out4[1] = s.by; out4[2] = s.red_x; out4[3] = s.red_y;
    return 0;
}
int emu_dit2_j0_max_stage(void) { return dit2_j0_trivial() ? kJ0MaxStage : -1; }
int emu_lazy_full_ok(int word_bytes, uint64_t q, int logn) {
    return word_bytes == 4 ? host::lazy_full_ok<uint32_t>(q, logn) : host::lazy_full_ok<uint64_t>(q, logn);
}
}
