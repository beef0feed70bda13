// Multi-modulus (RNS) engine: SURVEY.md section 8 row f3 (reports/final-report.tex:1811-1817 names RNS / FHE
// parameter sets as the step after the single 60-bit modulus).
//
//   * operands are [L][batch][N]: limb l holds the residues mod q_l; ONE launch multiplies all limbs
//     (polymul_rns_kernel, kernels.cuh: blockIdx.y picks the limb's tables out of the kernel parameters);
//   * the tables of all limbs -- what scripts/generate_twiddles.py:29-41 and generate_inverse_twiddles.py:48-61
//     write to rtl/*.hex for one modulus, in the orders the kernels read them, plus the Shoup companions -- are
//     GENERATED ON THE DEVICE by one kernel from (q_l, psi_l): plan creation for 16-32 limbs is L small host
//     scalar computations (inverses) and one launch, not L host table loops and uploads;
//   * scripts/find_psi.py:9-44 is tntt_find_psi (host side, like the reference's script).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/tntt.h"
#include "common.h"
#include "tables.h"

using namespace tntt;

namespace {

#define CUDA_TRY(expr)                                                                               \
    do {                                                                                             \
        cudaError_t e_ = (expr);                                                                     \
        if (e_ != cudaSuccess) return api_fail(TNTT_CUDA_ERROR, "%s: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

// ---------------------------------------------------------------------------------------------
// kernel shapes of the RNS engine: the default fused shape of each (word, n, reduction) class
// ---------------------------------------------------------------------------------------------
struct RnsVariant {
    const char *name;
    int word_bytes, logn, logr, red;
    cudaError_t (*launch)(const void *a, const void *b, void *c, size_t batch, const void *limbs, int nlimbs, cudaStream_t st);
    cudaError_t (*prepare)();
    cudaError_t (*attributes)(cudaFuncAttributes *attr, int *blocks_per_sm);
    // transform-domain kernels of the same shape: op 0 = forward spectrum, 1 = inverse spectrum, 2 = product with b spectra
    cudaError_t (*spectrum)(int op, const void *a, const void *b, void *c, size_t batch, size_t b_stride, const void *limbs,
                            int nlimbs, cudaStream_t st);
};

template <class C, int NA, int RED, int MINB, int STASH = 0> struct RnsInst {
    using W = typename C::W;
    static constexpr size_t SMEM = ((size_t)NA * C::TILE + (size_t)STASH * C::PPC * C::N) * sizeof(W);
    static cudaError_t launch(const void *a, const void *b, void *c, size_t batch, const void *limbs, int nlimbs, cudaStream_t st) {
        if (batch == 0 || nlimbs == 0) return cudaSuccess;
        const size_t ctas = (batch + C::PPC - 1) / C::PPC;
        polymul_rns_kernel<C, NA, RED, MINB, STASH><<<dim3((unsigned)ctas, (unsigned)nlimbs), C::THREADS, SMEM, st>>>(
            static_cast<const W *>(a), static_cast<const W *>(b), static_cast<W *>(c), batch,
            *static_cast<const RnsLimbs<W> *>(limbs));
        return cudaGetLastError();
    }
    // the transform-domain kernels hold one operand: one tile, the occupancy of the single-modulus spectrum kernels
    // (64-bit words: the unpadded tile, whose twiddle-group depth is the one tuned for one operand at 80 registers; the
    // spectrum order depends on R and P only, so it is the order of the single-modulus plans of the same size)
    using CS = Cfg<W, C::LOGN, C::LOGR, C::PPC, (sizeof(W) == 4 ? C::PAD : 0)>;
    static constexpr size_t SP_SMEM = (size_t)CS::TILE * sizeof(W);
    static constexpr int SP_MINB = (sizeof(W) == 4 && C::LOGR == 5) ? 2 : (NA == 2 ? MINB + 1 : MINB);
    static cudaError_t spectrum(int op, const void *a, const void *b, void *c, size_t batch, size_t b_stride, const void *limbs,
                                int nlimbs, cudaStream_t st) {
        if (batch == 0 || nlimbs == 0) return cudaSuccess;
        const dim3 grid((unsigned)((batch + C::PPC - 1) / C::PPC), (unsigned)nlimbs);
        const W *pa = static_cast<const W *>(a), *pb = static_cast<const W *>(b);
        W *pc = static_cast<W *>(c);
        const RnsLimbs<W> &L = *static_cast<const RnsLimbs<W> *>(limbs);
        if (op == 0) spectrum_rns_kernel<CS, RED, SP_MINB, 0><<<grid, C::THREADS, SP_SMEM, st>>>(pa, pb, pc, batch, b_stride, L);
        else if (op == 1) spectrum_rns_kernel<CS, RED, SP_MINB, 1><<<grid, C::THREADS, SP_SMEM, st>>>(pa, pb, pc, batch, b_stride, L);
        else spectrum_rns_kernel<CS, RED, SP_MINB, 2><<<grid, C::THREADS, SP_SMEM, st>>>(pa, pb, pc, batch, b_stride, L);
        return cudaGetLastError();
    }
    static cudaError_t prepare() {
        cudaError_t e = cudaFuncSetAttribute(polymul_rns_kernel<C, NA, RED, MINB, STASH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(spectrum_rns_kernel<CS, RED, SP_MINB, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SP_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(spectrum_rns_kernel<CS, RED, SP_MINB, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SP_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(spectrum_rns_kernel<CS, RED, SP_MINB, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SP_SMEM);
        return e;
    }
    static cudaError_t attributes(cudaFuncAttributes *attr, int *blocks_per_sm) {
        cudaError_t e = cudaFuncGetAttributes(attr, polymul_rns_kernel<C, NA, RED, MINB, STASH>);
        if (e != cudaSuccess) return e;
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, polymul_rns_kernel<C, NA, RED, MINB, STASH>, C::THREADS, SMEM);
    }
};
#define RNS_VARIANT(NAME, WT, LN, LR, PPC, PAD, NA, RED, MINB, ST)                                              \
    RnsVariant { NAME, (int)sizeof(WT), LN, LR, RED, &RnsInst<Cfg<WT, LN, LR, PPC, PAD>, NA, RED, MINB, ST>::launch, \
                 &RnsInst<Cfg<WT, LN, LR, PPC, PAD>, NA, RED, MINB, ST>::prepare,                                \
                 &RnsInst<Cfg<WT, LN, LR, PPC, PAD>, NA, RED, MINB, ST>::attributes,                             \
                 &RnsInst<Cfg<WT, LN, LR, PPC, PAD>, NA, RED, MINB, ST>::spectrum }

const RnsVariant kRnsVariants[] = {
    RNS_VARIANT("rns_u64_n12_r4_p1_a2_red1_b2_pad", uint64_t, 12, 4, 1, 1, 2, 1, 2, 0),
    RNS_VARIANT("rns_u64_n12_r4_p1_a2_red0_b2_pad", uint64_t, 12, 4, 1, 1, 2, 0, 2, 0),
    RNS_VARIANT("rns_u64_n10_r4_p4_a1_red1_b2", uint64_t, 10, 4, 4, 0, 1, 1, 2, 0),
    RNS_VARIANT("rns_u64_n10_r4_p4_a1_red0_b2", uint64_t, 10, 4, 4, 0, 1, 0, 2, 0),
    RNS_VARIANT("rns_u64_n8_r4_p16_a1_red1_b2", uint64_t, 8, 4, 16, 0, 1, 1, 2, 0),
    RNS_VARIANT("rns_u64_n8_r4_p16_a1_red0_b2", uint64_t, 8, 4, 16, 0, 1, 0, 2, 0),
    RNS_VARIANT("rns_u32_n12_r4_p1_a2_red0_b3_pad", uint32_t, 12, 4, 1, 1, 2, 0, 3, 0),
    RNS_VARIANT("rns_u32_n12_r4_p1_a2_red0_b3", uint32_t, 12, 4, 1, 0, 2, 0, 3, 0),     // alternatives: TNTT_RNS_PICK=1, 2, 3 (tools/rns_bench.py)
    RNS_VARIANT("rns_u32_n12_r4_p1_a2_red0_b2", uint32_t, 12, 4, 1, 0, 2, 0, 2, 0),
    RNS_VARIANT("rns_u32_n12_r4_p1_a1_red0_b4", uint32_t, 12, 4, 1, 0, 1, 0, 4, 0),
    RNS_VARIANT("rns_u32_n10_r5_p8_a2_red0_b2_pad", uint32_t, 10, 5, 8, 1, 2, 0, 2, 0),
    RNS_VARIANT("rns_u32_n8_r4_p16_a2_red0_b3_pad", uint32_t, 8, 4, 16, 1, 2, 0, 3, 0),
};

// ---------------------------------------------------------------------------------------------
// device-side table generation
// ---------------------------------------------------------------------------------------------
typedef unsigned __int128 u128;

struct GenLimb {          // one limb's inputs and output pointers (device array)
    uint64_t q, psi, psi_inv, omega_inv, post_scale, n_inv;
    void *fwd_pyr, *fwd_last, *inv_pyr, *post, *post_untwist;
};

__device__ __forceinline__ uint64_t d_mulmod(uint64_t a, uint64_t b, uint64_t q) { return (uint64_t)(((u128)a * b) % q); }
__device__ uint64_t d_powmod(uint64_t b, uint64_t e, uint64_t q) {
    uint64_t r = 1 % q;
    for (; e; e >>= 1, b = d_mulmod(b, b, q))
        if (e & 1) r = d_mulmod(r, b, q);
    return r;
}
template <typename W> __device__ __forceinline__ Tw<W> d_make_tw(uint64_t w, uint64_t q) {   // host::make_tw
    return Tw<W>{(W)w, (W)((((u128)w) << WordTraits<W>::BITS) / q)};
}
__device__ __forceinline__ uint32_t d_bitrev(uint32_t v, int bits) { return bits ? __brev(v) >> (32 - bits) : 0; }

// grid (ceil(max(n, (R-1)P) / 256), L).  Entry k of every table of limb blockIdx.y; definitions = tables.h:
//   fwd_pyr[k]  = psi^bitrev(k)                           (host::fwd_pyramid)
//   inv_pyr[2^b + j] = omega^-(j n / 2^(b+1)), [0] = 1    (host::dit_pyramid)
//   post[k] = post_scale psi^-k, post_untwist[k] = n^-1 psi^-k   (host::scaled_powers)
//   fwd_last[slot P + tid] = fwd_pyr[2^s + (tid << l) + g]       (host::fwd_last_table)
template <typename W>
__global__ void rns_tables_kernel(const GenLimb *limbs, int logn, int logr) {
    const GenLimb g = limbs[blockIdx.y];
    const uint32_t n = 1u << logn, k = blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t q = g.q;
    if (k < n) {
        static_cast<Tw<W> *>(g.fwd_pyr)[k] = d_make_tw<W>(d_powmod(g.psi, d_bitrev(k, logn), q), q);
        uint64_t v = 1 % q;
        if (k > 0) {
            const int b = 31 - __clz(k);
            const uint32_t j = k - (1u << b);
            v = d_powmod(g.omega_inv, (uint64_t)j << (logn - 1 - b), q);
        }
        static_cast<Tw<W> *>(g.inv_pyr)[k] = d_make_tw<W>(v, q);
        const uint64_t pk = d_powmod(g.psi_inv, k, q);
        static_cast<Tw<W> *>(g.post)[k] = d_make_tw<W>(d_mulmod(g.post_scale % q, pk, q), q);
        static_cast<Tw<W> *>(g.post_untwist)[k] = d_make_tw<W>(d_mulmod(g.n_inv, pk, q), q);
    }
    const int R = 1 << logr, P = 1 << (logn - logr);
    if (k < (uint32_t)((R - 1) * P)) {
        const int npass = (logn + logr - 1) / logr, bhi = logn - (npass - 1) * logr;
        const int slot = (int)(k / P), tid = (int)(k % P);
        const int l = 31 - __clz(slot + 1), gi = slot + 1 - (1 << l), kb = logr - 1 - l;
        Tw<W> e{0, 0};
        if (kb < bhi) {
            const int s = logn - 1 - kb;
            const uint32_t src = (1u << s) + ((uint32_t)tid << l) + gi;
            e = d_make_tw<W>(d_powmod(g.psi, d_bitrev(src, logn), q), q);
        }
        static_cast<Tw<W> *>(g.fwd_last)[k] = e;
    }
}

template <typename W> size_t tw_bytes(size_t entries) { return entries * sizeof(Tw<W>); }

}  // namespace

struct tntt_rns_plan {
    int device = 0, limbs = 0, word_bytes = 0, logn = 0, red = 0;
    uint32_t n = 0;
    std::vector<uint64_t> q, psi;
    const RnsVariant *var = nullptr;
    void *slab = nullptr;               // all tables of all limbs
    size_t slab_bytes = 0;
    std::vector<GenLimb> gen;           // host copy (device pointers inside)
    // launch descriptors, kRnsMaxLimbs limbs per group, by value into the kernel parameters
    std::vector<RnsLimbs<uint64_t>> groups64;
    std::vector<RnsLimbs<uint32_t>> groups32;
};

namespace {

template <typename W> void fill_group(tntt_rns_plan *p, std::vector<RnsLimbs<W>> &groups) {
    const int ngroups = (p->limbs + kRnsMaxLimbs - 1) / kRnsMaxLimbs;
    groups.assign(ngroups, RnsLimbs<W>{});
    for (int l = 0; l < p->limbs; ++l) {
        RnsLimb<W> &L = groups[l / kRnsMaxLimbs].limb[l % kRnsMaxLimbs];
        const GenLimb &g = p->gen[l];
        L.tb.fwd_pyr = static_cast<const Tw<W> *>(g.fwd_pyr);
        L.tb.fwd_last = static_cast<const Tw<W> *>(g.fwd_last);
        L.tb.post = static_cast<const Tw<W> *>(g.post);
        L.tb.inv.pyr = static_cast<const Tw<W> *>(g.inv_pyr);
        L.untwist = static_cast<const Tw<W> *>(g.post_untwist);
        L.mod = host::make_mod<W>(g.q, p->logn);
    }
}
// the first MAX_R entries of the pyramids travel by value (uniform twiddles of the first passes)
template <typename W> int fetch_heads(tntt_rns_plan *p, std::vector<RnsLimbs<W>> &groups) {
    for (int l = 0; l < p->limbs; ++l) {
        RnsLimb<W> &L = groups[l / kRnsMaxLimbs].limb[l % kRnsMaxLimbs];
        const size_t head = (size_t)(p->n < (uint32_t)MAX_R ? p->n : MAX_R) * sizeof(Tw<W>);
        CUDA_TRY(cudaMemcpy(L.tb.fwd_head, p->gen[l].fwd_pyr, head, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(L.tb.inv.head, p->gen[l].inv_pyr, head, cudaMemcpyDeviceToHost));
    }
    return TNTT_OK;
}

template <typename W> int check_tables(const tntt_rns_plan *p, int limb) {
    const GenLimb &g = p->gen[limb];
    const uint32_t n = p->n;
    const int logr = p->var->logr;
    auto same = [&](const std::vector<Tw<W>> &want, const void *dev, const char *what) -> int {
        std::vector<Tw<W>> got(want.size());
        CUDA_TRY(cudaMemcpy(got.data(), dev, want.size() * sizeof(Tw<W>), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < want.size(); ++i)
            if (got[i].w != want[i].w || got[i].wp != want[i].wp)
                return api_fail(TNTT_BAD_ARG, "limb %d: device-generated %s[%zu] differs from the host generator", limb, what, i);
        return TNTT_OK;
    };
    const std::vector<Tw<W>> fwd = host::fwd_pyramid<W>(g.psi, n, g.q);
    int rc = same(fwd, g.fwd_pyr, "fwd_pyr");
    if (!rc) rc = same(host::fwd_last_table<W>(fwd, p->logn, logr), g.fwd_last, "fwd_last");
    if (!rc) rc = same(host::dit_pyramid<W>(g.omega_inv, n, g.q), g.inv_pyr, "inv_pyr");
    if (!rc) rc = same(host::scaled_powers<W>(g.psi_inv, g.post_scale, n, g.q), g.post, "post");
    if (!rc) rc = same(host::scaled_powers<W>(g.psi_inv, g.n_inv, n, g.q), g.post_untwist, "post_untwist");
    return rc;
}

struct DevGuard {
    int prev = -1;
    bool ok = true;
    explicit DevGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

}  // namespace

extern "C" {
#pragma GCC visibility push(default)

int tntt_find_psi(uint32_t n, uint64_t q, uint64_t max_search, uint64_t *psi) {
    if (!psi) return api_fail(TNTT_BAD_ARG, "psi is null");
    if (n < 2 || (n & (n - 1))) return api_fail(TNTT_UNSUPPORTED_N, "n=%u must be a power of two", n);
    if (q < 3 || !host::is_prime(q)) return api_fail(TNTT_UNSUPPORTED_Q, "q=%llu is not an odd prime", (unsigned long long)q);
    if ((q - 1) % (2ull * n)) return api_fail(TNTT_BAD_ROOT, "q=%llu is not 1 mod 2n: no primitive 2n-th root exists", (unsigned long long)q);
    // scripts/find_psi.py:29-30: the smallest integer in [2, max_search) with psi^n = -1 (then psi^2n = 1 follows)
    for (uint64_t c = 2; c < max_search && c < q; ++c)
        if (host::powmod(c, n, q) == q - 1) { *psi = c; return TNTT_OK; }
    // none that small (the reference gives up here): g^((q-1)/2n) for the first g that yields order exactly 2n
    const uint64_t e = (q - 1) / (2ull * n);
    for (uint64_t g = 2; g < q && g < (1ull << 20); ++g) {
        const uint64_t c = host::powmod(g, e, q);
        if (host::powmod(c, n, q) == q - 1) { *psi = c; return 1; }   // 1: found outside the reference's search range
    }
    return api_fail(TNTT_BAD_ROOT, "no primitive 2n-th root found");
}

int tntt_rns_plan_create(tntt_rns_plan **out, int device, uint32_t n, const uint64_t *q, const uint64_t *psi, int limbs) {
    if (!out) return api_fail(TNTT_BAD_ARG, "out is null");
    *out = nullptr;
    if (!q || !psi || limbs < 1 || limbs > 4096) return api_fail(TNTT_BAD_ARG, "need 1..4096 limbs and their (q, psi)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return api_fail(TNTT_NO_DEVICE, "no CUDA device visible; libtntt has no CPU path");
    }
    if (device < 0 || device >= ndev) return api_fail(TNTT_BAD_ARG, "device %d out of range (0..%d)", device, ndev - 1);
    if (n < 2 || n > 65536 || (n & (n - 1))) return api_fail(TNTT_UNSUPPORTED_N, "n=%u must be a power of two in [2, 65536]", n);
    const int logn = host::ilog2(n);
    int word = 0, red = 0;
    for (int l = 0; l < limbs; ++l) {
        const uint64_t ql = q[l];
        if (ql < 3 || !(ql & 1) || ql >= (1ull << 60) || !host::is_prime(ql))
            return api_fail(TNTT_UNSUPPORTED_Q, "limb %d: q=%llu must be an odd prime below 2^60", l, (unsigned long long)ql);
        for (int m = 0; m < l; ++m)
            if (q[m] == ql) return api_fail(TNTT_UNSUPPORTED_Q, "limb %d repeats the modulus of limb %d (CRT needs distinct primes)", l, m);
        if (psi[l] >= ql || !host::is_primitive_2n_root(psi[l], n, ql))
            return api_fail(TNTT_BAD_ROOT, "limb %d: psi=%llu: psi^%u != -1 mod q", l, (unsigned long long)psi[l], n);
        const int w = host::lazy_full_ok<uint32_t>(ql, logn) ? 4 : 8;
        const int r = (w == 8 && !host::lazy_full_ok<uint64_t>(ql, logn)) ? 1 : 0;
        if (l == 0) { word = w; red = r; }
        else if (w != word) return api_fail(TNTT_UNSUPPORTED_Q, "limb %d needs %d-byte words, limb 0 %d-byte ones: one word size per RNS plan", l, w, word);
        else if (r > red) red = r;      // the lazily reducing kernels serve every modulus of their word size
    }
    const RnsVariant *var = nullptr;
    int pick = 0;   // the first matching shape is the measured default; TNTT_RNS_PICK=k selects the k-th match (benchmarking)
    if (const char *env = getenv("TNTT_RNS_PICK")) pick = atoi(env);
    for (const RnsVariant &v : kRnsVariants)
        if (v.word_bytes == word && v.logn == logn && v.red == red) {
            var = &v;
            if (pick-- <= 0) break;
        }
    if (!var) return api_fail(TNTT_UNSUPPORTED_N, "no multi-modulus kernel for n=%u with %d-byte words (n in {256, 1024, 4096})", n, word);

    DevGuard dg(device);
    if (!dg.ok) return api_fail(TNTT_CUDA_ERROR, "cudaSetDevice(%d) failed", device);
    tntt_rns_plan *p = new tntt_rns_plan();
    p->device = device; p->limbs = limbs; p->word_bytes = word; p->logn = logn; p->red = red; p->n = n; p->var = var;
    p->q.assign(q, q + limbs);
    p->psi.assign(psi, psi + limbs);
    const size_t tw = word == 4 ? sizeof(Tw<uint32_t>) : sizeof(Tw<uint64_t>);
    const size_t R = (size_t)1 << var->logr, P = (size_t)n >> var->logr;
    const size_t per_table = (size_t)(n > 1 ? n : 2) * tw, last_bytes = (R - 1) * P * tw;
    const size_t per_limb = 4 * per_table + last_bytes;
    p->slab_bytes = per_limb * limbs;
    cudaError_t e = cudaMalloc(&p->slab, p->slab_bytes);
    if (e != cudaSuccess) { delete p; return api_fail(TNTT_CUDA_ERROR, "cudaMalloc(%zu): %s", per_limb * limbs, cudaGetErrorString(e)); }
    p->gen.resize(limbs);
    const int bits = word * 8;
    for (int l = 0; l < limbs; ++l) {
        GenLimb &g = p->gen[l];
        unsigned char *base = static_cast<unsigned char *>(p->slab) + per_limb * l;
        g.q = q[l]; g.psi = psi[l];
        g.psi_inv = host::modinv(psi[l], q[l]);
        g.omega_inv = host::modinv(host::mulmod(psi[l], psi[l], q[l]), q[l]);
        g.n_inv = host::modinv(n % q[l], q[l]);
        // the Montgomery pointwise product leaves 2^-BITS for the store table to undo (PolymulTables::post)
        g.post_scale = host::mulmod(g.n_inv, (uint64_t)((((host::u128)1) << bits) % q[l]), q[l]);
        g.fwd_pyr = base; g.inv_pyr = base + per_table; g.post = base + 2 * per_table; g.post_untwist = base + 3 * per_table;
        g.fwd_last = base + 4 * per_table;
    }
    GenLimb *dgen = nullptr;
    auto bail = [&](int rc) { if (dgen) cudaFree(dgen); tntt_rns_plan_destroy(p); return rc; };
    if ((e = cudaMalloc(&dgen, sizeof(GenLimb) * limbs)) != cudaSuccess) return bail(api_fail(TNTT_CUDA_ERROR, "cudaMalloc: %s", cudaGetErrorString(e)));
    if ((e = cudaMemcpy(dgen, p->gen.data(), sizeof(GenLimb) * limbs, cudaMemcpyHostToDevice)) != cudaSuccess)
        return bail(api_fail(TNTT_CUDA_ERROR, "cudaMemcpy: %s", cudaGetErrorString(e)));
    const size_t entries = n > (R - 1) * P ? n : (R - 1) * P;
    const dim3 grid((unsigned)((entries + 255) / 256), (unsigned)limbs);
    if (word == 4) rns_tables_kernel<uint32_t><<<grid, 256>>>(dgen, logn, var->logr);
    else rns_tables_kernel<uint64_t><<<grid, 256>>>(dgen, logn, var->logr);
    if ((e = cudaGetLastError()) != cudaSuccess || (e = cudaDeviceSynchronize()) != cudaSuccess)
        return bail(api_fail(TNTT_CUDA_ERROR, "table generation kernel: %s", cudaGetErrorString(e)));
    cudaFree(dgen);
    dgen = nullptr;
    int rc;
    if (word == 4) { fill_group<uint32_t>(p, p->groups32); rc = fetch_heads<uint32_t>(p, p->groups32); }
    else { fill_group<uint64_t>(p, p->groups64); rc = fetch_heads<uint64_t>(p, p->groups64); }
    if (rc) return bail(rc);
    if ((e = var->prepare()) != cudaSuccess) return bail(api_fail(TNTT_CUDA_ERROR, "prepare %s: %s", var->name, cudaGetErrorString(e)));
    *out = p;
    return TNTT_OK;
}

void tntt_rns_plan_destroy(tntt_rns_plan *p) {
    if (!p) return;
    DevGuard dg(p->device);
    if (p->slab) cudaFree(p->slab);
    delete p;
}

int tntt_rns_plan_limbs(const tntt_rns_plan *p) { return p ? p->limbs : 0; }
int tntt_rns_plan_word_bytes(const tntt_rns_plan *p) { return p ? p->word_bytes : 0; }
const char *tntt_rns_plan_kernel(const tntt_rns_plan *p) { return p ? p->var->name : ""; }
size_t tntt_rns_plan_table_bytes(const tntt_rns_plan *p) { return p ? p->slab_bytes : 0; }

int tntt_rns_plan_check_tables(const tntt_rns_plan *p, int limb) {
    if (!p || limb < 0 || limb >= p->limbs) return api_fail(TNTT_BAD_ARG, "bad plan or limb index");
    DevGuard dg(p->device);
    return p->word_bytes == 4 ? check_tables<uint32_t>(p, limb) : check_tables<uint64_t>(p, limb);
}

int tntt_rns_kernel_attributes(const tntt_rns_plan *p, int *regs, size_t *local_bytes, int *ctas_per_sm) {
    if (!p) return api_fail(TNTT_BAD_ARG, "plan is null");
    DevGuard dg(p->device);
    cudaFuncAttributes attr{};
    int occ = 0;
    CUDA_TRY(p->var->attributes(&attr, &occ));
    if (regs) *regs = attr.numRegs;
    if (local_bytes) *local_bytes = attr.localSizeBytes;
    if (ctas_per_sm) *ctas_per_sm = occ;
    return TNTT_OK;
}

int tntt_rns_polymul(const tntt_rns_plan *p, const void *a, const void *b, void *c, size_t batch, void *stream) {
    if (!p) return api_fail(TNTT_BAD_ARG, "plan is null");
    if (batch == 0) return TNTT_OK;
    if (!a || !b || !c) return api_fail(TNTT_BAD_ARG, "null data pointer");
    if (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15) return api_fail(TNTT_BAD_ARG, "data pointers must be 16-byte aligned");
    DevGuard dg(p->device);
    const size_t limb_bytes = batch * p->n * (size_t)p->word_bytes;
    const int ngroups = (p->limbs + kRnsMaxLimbs - 1) / kRnsMaxLimbs;
    for (int g = 0; g < ngroups; ++g) {
        const int first = g * kRnsMaxLimbs, count = p->limbs - first < kRnsMaxLimbs ? p->limbs - first : kRnsMaxLimbs;
        const void *limbs = p->word_bytes == 4 ? (const void *)&p->groups32[g] : (const void *)&p->groups64[g];
        const size_t off = limb_bytes * first;
        CUDA_TRY(p->var->launch(static_cast<const char *>(a) + off, static_cast<const char *>(b) + off,
                                static_cast<char *>(c) + off, batch, limbs, count, (cudaStream_t)stream));
    }
    return TNTT_OK;
}

namespace {
int rns_io_check(const tntt_rns_plan *p, const void *a, const void *b, const void *c) {
    if (!p) return api_fail(TNTT_BAD_ARG, "plan is null");
    if (!a || !b || !c) return api_fail(TNTT_BAD_ARG, "null data pointer");
    if (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15) return api_fail(TNTT_BAD_ARG, "data pointers must be 16-byte aligned");
    return TNTT_OK;
}
// op 0 / 1 / 2 of RnsVariant::spectrum, one launch per kRnsMaxLimbs limbs
int rns_spectrum(const tntt_rns_plan *p, int op, const void *a, const void *b, void *c, size_t batch, size_t b_rows, void *stream) {
    if (batch == 0) return p ? TNTT_OK : api_fail(TNTT_BAD_ARG, "plan is null");
    int rc = rns_io_check(p, a, op == 2 ? b : a, c);
    if (rc) return rc;
    if (op == 2 && b_rows != 1 && b_rows != batch) return api_fail(TNTT_BAD_ARG, "b_rows must be 1 (one spectrum per limb) or the batch size");
    DevGuard dg(p->device);
    const size_t row_bytes = (size_t)p->n * p->word_bytes;
    const int ngroups = (p->limbs + kRnsMaxLimbs - 1) / kRnsMaxLimbs;
    for (int g = 0; g < ngroups; ++g) {
        const int first = g * kRnsMaxLimbs, count = p->limbs - first < kRnsMaxLimbs ? p->limbs - first : kRnsMaxLimbs;
        const void *limbs = p->word_bytes == 4 ? (const void *)&p->groups32[g] : (const void *)&p->groups64[g];
        const size_t off = row_bytes * batch * first, boff = op == 2 ? row_bytes * b_rows * first : 0;
        CUDA_TRY(p->var->spectrum(op, static_cast<const char *>(a) + off, b ? static_cast<const char *>(b) + boff : nullptr,
                                  static_cast<char *>(c) + off, batch, (op == 2 && b_rows == batch) ? p->n : 0, limbs, count,
                                  (cudaStream_t)stream));
    }
    return TNTT_OK;
}
}  // namespace

int tntt_rns_spectrum_forward(const tntt_rns_plan *p, const void *in, void *out, size_t batch, void *stream) {
    return rns_spectrum(p, 0, in, nullptr, out, batch, 0, stream);
}
int tntt_rns_spectrum_inverse(const tntt_rns_plan *p, const void *in, void *out, size_t batch, void *stream) {
    return rns_spectrum(p, 1, in, nullptr, out, batch, 0, stream);
}
int tntt_rns_polymul_spectrum(const tntt_rns_plan *p, const void *a, const void *b_spectrum, void *c, size_t batch, size_t b_rows,
                              void *stream) {
    return rns_spectrum(p, 2, a, b_spectrum, c, batch, b_rows, stream);
}
int tntt_rns_pointwise(const tntt_rns_plan *p, const void *a, const void *b, void *c, size_t batch, void *stream) {
    if (batch == 0) return p ? TNTT_OK : api_fail(TNTT_BAD_ARG, "plan is null");
    int rc = rns_io_check(p, a, b, c);
    if (rc) return rc;
    DevGuard dg(p->device);
    const size_t count = batch * p->n, row_bytes = (size_t)p->n * p->word_bytes;
    const int ngroups = (p->limbs + kRnsMaxLimbs - 1) / kRnsMaxLimbs;
    for (int g = 0; g < ngroups; ++g) {
        const int first = g * kRnsMaxLimbs, nl = p->limbs - first < kRnsMaxLimbs ? p->limbs - first : kRnsMaxLimbs;
        const size_t off = row_bytes * batch * first;
        const size_t want = (count + 255) / 256;
        const dim3 grid((unsigned)(want < 148u * 16u ? (want ? want : 1) : 148u * 16u), (unsigned)nl);
        if (p->word_bytes == 4)
            pointwise_rns_kernel<uint32_t><<<grid, 256, 0, (cudaStream_t)stream>>>(
                (const uint32_t *)((const char *)a + off), (const uint32_t *)((const char *)b + off), (uint32_t *)((char *)c + off), count, p->groups32[g]);
        else
            pointwise_rns_kernel<uint64_t><<<grid, 256, 0, (cudaStream_t)stream>>>(
                (const uint64_t *)((const char *)a + off), (const uint64_t *)((const char *)b + off), (uint64_t *)((char *)c + off), count, p->groups64[g]);
        CUDA_TRY(cudaGetLastError());
    }
    return TNTT_OK;
}

#pragma GCC visibility pop
}
