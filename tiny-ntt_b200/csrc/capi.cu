// C ABI of libtntt.so (include/tntt.h): plans, table upload, kernel dispatch, host pipeline.
// No CPU compute path exists here: every entry point that produces coefficients launches a
// CUDA kernel or fails.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/tntt.h"
#include "common.h"
#include "tables.h"

using namespace tntt;

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
}  // namespace
// the same, for the other translation units of the library (rns.cu, multi.cu)
int tntt::api_fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
namespace {
#define CUDA_TRY(expr)                                                                           \
    do {                                                                                         \
        cudaError_t e_ = (expr);                                                                 \
        if (e_ != cudaSuccess) return fail(TNTT_CUDA_ERROR, "%s: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

// all fused variants, in a stable order
std::vector<PolymulVariant> &all_variants() {
    static std::vector<PolymulVariant> v = [] {
        std::vector<PolymulVariant> out;
        int c = 0;
        const PolymulVariant *p = polymul_variants_u32(&c);
        out.insert(out.end(), p, p + c);
        p = polymul_variants_u64(&c);
        out.insert(out.end(), p, p + c);
        p = polymul_variants_u64b(&c);
        out.insert(out.end(), p, p + c);
        return out;
    }();
    return v;
}

struct DeviceSetter {
    int prev = -1;
    bool ok = true;
    explicit DeviceSetter(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceSetter() { if (prev >= 0) cudaSetDevice(prev); }
};

template <typename T> cudaError_t upload(const std::vector<T> &h, void **d) {
    cudaError_t e = cudaMalloc(d, h.size() * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
}

constexpr int kMaxLogR = 6;

}  // namespace

struct tntt_plan {
    tntt_plan_info info{};
    Mod<uint32_t> mod32{};
    Mod<uint64_t> mod64{};
    std::vector<uint64_t> psi_pow, psi_inv_pow;  // host copies (hex export)
    // device tables (Tw<W> unless noted)
    void *fwd_pyr = nullptr, *inv_pyr = nullptr, *post_mont = nullptr;  // fused polymul
    void *fwd_last[kMaxLogR + 1] = {};                                 // per LOGR
    void *cyc_fwd_pyr = nullptr;                                       // DIT pyramid of omega (cg_ntt)
    void *cyc_ct_pyr = nullptr, *cyc_ct_last[kMaxLogR + 1] = {};       // cyclic Cooley-Tukey pyramid of omega (+ transposed last pass)
    void *pre_twist = nullptr, *post_untwist = nullptr;                // psi^i ; psi^-i N^-1
    void *omega_pow = nullptr, *omega_inv_pow = nullptr;               // plain W[n]: literal CG stages
    Tw<uint64_t> one_tw{}, ninv_tw{};
    // first MAX_R entries of the pyramids, passed by value in the kernel parameters
    Tw<uint32_t> head32[4][MAX_R] = {};   // [0] fwd_pyr, [1] inv_pyr, [2] cyc_fwd_pyr, [3] cyc_ct_pyr
    Tw<uint64_t> head64[4][MAX_R] = {};
    const TransformVariant *xform = nullptr;
    const SpectrumVariant *spectrum = nullptr;
    // host pipeline
    std::mutex pipe_mu;
    static constexpr int kSlots = 4;
    cudaStream_t pipe_stream[kSlots] = {};
    void *pipe_buf[kSlots][3] = {};
    size_t pipe_rows = 0;

    const void *mod() const { return info.word_bytes == 4 ? (const void *)&mod32 : (const void *)&mod64; }
};

namespace {

template <typename W> int build_tables(tntt_plan *p) {
    const uint32_t n = p->info.n;
    const uint64_t q = p->info.q;
    const int logn = (int)p->info.logn;
    constexpr int BITS = WordTraits<W>::BITS;
    auto keep_head = [&](const std::vector<Tw<W>> &v, int which) {
        for (size_t i = 0; i < (size_t)MAX_R && i < v.size(); ++i) {
            if constexpr (sizeof(W) == 4) p->head32[which][i] = v[i];
            else p->head64[which][i] = v[i];
        }
    };
    if (p->info.omega_is_primitive) {
        const std::vector<Tw<W>> cf = host::dit_pyramid<W>(p->info.omega, n, q), ci = host::dit_pyramid<W>(p->info.omega_inv, n, q);
        CUDA_TRY(upload(cf, &p->cyc_fwd_pyr));
        CUDA_TRY(upload(ci, &p->inv_pyr));
        keep_head(cf, 2);
        keep_head(ci, 1);
        // cyclic Cooley-Tukey tables for the shapes of spectrum.cu (natural-order cg_ntt without a bit-reversal pass)
        int sc = 0;
        const SpectrumVariant *sv = spectrum_variants(&sc);
        for (int i = 0; i < sc; ++i)
            if (sv[i].word_bytes == (int)sizeof(W) && sv[i].logn == logn && !p->cyc_ct_last[sv[i].logr]) {
                const std::vector<Tw<W>> ct = host::fwd_pyramid_cyclic<W>(p->info.omega, n, q);
                if (!p->cyc_ct_pyr) { CUDA_TRY(upload(ct, &p->cyc_ct_pyr)); keep_head(ct, 3); }
                CUDA_TRY(upload(host::fwd_last_table<W>(ct, logn, sv[i].logr), &p->cyc_ct_last[sv[i].logr]));
            }
    }
    {   // natural power tables for the literal constant-geometry stages (any omega)
        std::vector<uint64_t> f = host::powers(p->info.omega, n, q), b = host::powers(p->info.omega_inv, n, q);
        std::vector<W> fw(f.begin(), f.end()), bw(b.begin(), b.end());
        CUDA_TRY(upload(fw, &p->omega_pow));
        CUDA_TRY(upload(bw, &p->omega_inv_pow));
    }
    if (p->info.has_psi && !p->info.literal_only) {   // tables of the fused kernels
        std::vector<Tw<W>> fwd = host::fwd_pyramid<W>(p->info.psi, n, q);
        CUDA_TRY(upload(fwd, &p->fwd_pyr));
        keep_head(fwd, 0);
        for (const PolymulVariant &v : all_variants())
            if (v.word_bytes == (int)sizeof(W) && v.logn == logn && !p->fwd_last[v.logr])
                CUDA_TRY(upload(host::fwd_last_table<W>(fwd, logn, v.logr), &p->fwd_last[v.logr]));
        const uint64_t r_mod_q = (uint64_t)((((host::u128)1) << BITS) % q);
        CUDA_TRY(upload(host::scaled_powers<W>(p->info.psi_inv, host::mulmod(p->info.n_inv, r_mod_q, q), n, q),
                        &p->post_mont));
    }
    if (p->info.has_psi) {   // the twist tables of cg_ntt.py:82-83,92 (any modulus)
        CUDA_TRY(upload(host::scaled_powers<W>(p->info.psi, 1, n, q), &p->pre_twist));
        CUDA_TRY(upload(host::scaled_powers<W>(p->info.psi_inv, p->info.n_inv, n, q), &p->post_untwist));
    }
    return TNTT_OK;
}

int choose_default_variant(const tntt_plan *p) {
    // preference order measured on B200 (profiles/): first match wins
    static const char *prefer[] = {
        "u64_n12_r4_p1_a2_red2_b2_s0_t0_pad", "u64_n12_r4_p1_a2_red1_b2_s0_t0_pad",
        "u64_n12_r4_p1_a1_red1_b3_s1_t0", "u64_n12_r4_p1_a2_red1_b2_s0_t0", "u64_n12_r4_p1_a1_red0_b2_s0_t0",
        "u32_n12_r4_p1_a2_red0_b3_s0_t0_pad", "u32_n10_r5_p8_a2_red0_b2_s0_t0_pad", "u32_n8_r4_p16_a2_red0_b3_s0_t0_pad",
        "u32_n12_r4_p1_a2_red0_b3_s0_t0", "u32_n10_r5_p8_a2_red0_b2_s0_t0", "u32_n8_r4_p16_a2_red0_b3_s0_t0",
    };
    const std::vector<PolymulVariant> &vs = all_variants();
    for (const char *name : prefer)
        for (size_t i = 0; i < vs.size(); ++i)
            if (!strcmp(vs[i].name, name) && tntt_variant_matches(p, (int)i)) return (int)i;
    for (size_t i = 0; i < vs.size(); ++i)
        if (!vs[i].cluster && tntt_variant_matches(p, (int)i)) return (int)i;
    // rows too long for one CTA (N = 16384, 32768): the cluster kernel is the only fused shape
    for (size_t i = 0; i < vs.size(); ++i)
        if (vs[i].cluster && tntt_variant_matches(p, (int)i)) {
            cudaFuncAttributes attr{};
            int clusters = 0;
            if (vs[i].attributes(&attr, &clusters) == cudaSuccess && clusters > 0) return (int)i;
            cudaGetLastError();
        }
    return -1;
}

int find_variant(const tntt_plan *p, const char *name) {
    const std::vector<PolymulVariant> &vs = all_variants();
    for (size_t i = 0; i < vs.size(); ++i)
        if (!strcmp(vs[i].name, name) && tntt_variant_matches(p, (int)i)) return (int)i;
    return -1;
}
// Small batches (measured on B200, profiles/r01_small_batch.jsonl): with fewer rows than SMs a row's latency is
// what counts.  Up to SMs/4 rows every row gets a cluster of 4 SMs; up to one row per SM the one-CTA-per-SM
// shape (all 64K registers, no stash tile) beats the three-CTAs-per-SM throughput shape.
void choose_small_batch_variants(tntt_plan *p) {
    tntt_plan_info &I = p->info;
    I.cluster_variant = I.small_variant = -1;
    I.cluster_batch_max = I.small_batch_max = 0;
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, I.device) != cudaSuccess || sms <= 0) return;
    const std::vector<PolymulVariant> &vs = all_variants();
    for (size_t i = 0; i < vs.size(); ++i)
        // 64-bit words only: the 32-bit rows are short enough that the cluster barriers eat the gain (measured)
        if (vs[i].cluster > 0 && vs[i].word_bytes == 8 && tntt_variant_matches(p, (int)i)) {
            cudaFuncAttributes attr{};
            int clusters = 0;
            if (vs[i].attributes(&attr, &clusters) != cudaSuccess || clusters <= 0) { cudaGetLastError(); continue; }
            I.cluster_variant = (int)i;
            I.cluster_batch_max = sms / vs[i].cluster < clusters ? sms / vs[i].cluster : clusters;
            break;
        }
    static const char *small[] = {"u64_n12_r3_p1_a1_red1_b1_s0_t0"};
    for (const char *name : small) {
        const int v = find_variant(p, name);
        if (v >= 0) { I.small_variant = v; I.small_batch_max = sms; break; }
    }
}

int create_plan(tntt_plan **out, int device, uint32_t n, uint64_t q, uint64_t root, int root_is_psi) {
    if (!out) return fail(TNTT_BAD_ARG, "out is null");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(TNTT_NO_DEVICE, "no CUDA device visible; libtntt has no CPU path");
    }
    if (device < 0 || device >= ndev) return fail(TNTT_BAD_ARG, "device %d out of range (0..%d)", device, ndev - 1);
    if (n < 2 || n > 65536 || (n & (n - 1))) return fail(TNTT_UNSUPPORTED_N, "n=%u must be a power of two in [2, 65536]", n);
    if (q < 2 || q >= (1ull << 60)) return fail(TNTT_UNSUPPORTED_Q, "q=%llu must be in [2, 2^60)", (unsigned long long)q);
    if (root >= q) return fail(TNTT_BAD_ROOT, "root must be reduced mod q");
    // new_reference/cg_ntt.py:29-92 never looks at its modulus or roots: `% modulus` after every operation and
    // modinv = pow(a, q - 2, q) give well-defined numbers for ANY integers, whether or not they are a transform of
    // anything.  The fast kernels (Shoup / Montgomery tables, merged-psi passes) need an odd prime q and primitive
    // roots; everything else -- even or composite q, psi with psi^n != -1, omega = 0 -- gets a LITERAL plan that runs
    // the reference's stage schedule as it is written (tntt_cg_stage kernels, Barrett products, the same Fermat-style
    // "inverses"), so the drop-in returns the reference's numbers there too.
    const bool field = q >= 3 && (q & 1) && host::is_prime(q);
    const bool literal = !field || (root_is_psi ? !host::is_primitive_2n_root(root, n, q) : root == 0);

    DeviceSetter ds(device);
    if (!ds.ok) return fail(TNTT_CUDA_ERROR, "cudaSetDevice(%d) failed", device);

    tntt_plan *p = new tntt_plan();
    tntt_plan_info &I = p->info;
    I.n = n;
    I.logn = (uint32_t)host::ilog2(n);
    I.q = q;
    I.device = device;
    I.has_psi = root_is_psi ? 1 : 0;
    I.psi = root_is_psi ? root : 0;
    I.psi_inv = root_is_psi ? host::modinv(root, q) : 0;
    I.omega = root_is_psi ? host::mulmod(root, root, q) : root;
    I.omega_inv = host::modinv(I.omega, q);
    I.omega_is_primitive = (!literal && host::is_primitive_n_root(I.omega, n, q)) ? 1 : 0;
    I.literal_only = literal ? 1 : 0;
    I.n_inv = host::modinv(n % q, q);
    // uint32 coefficients when the whole transform fits the lazy 32-bit range, else uint64
    I.word_bytes = host::lazy_full_ok<uint32_t>(q, (int)I.logn) ? 4 : 8;
    I.lazy_reduce = (I.word_bytes == 8 && !host::lazy_full_ok<uint64_t>(q, (int)I.logn)) ? 1 : 0;
    I.solinas = (q == kSolinasQ) ? 1 : 0;
    p->mod32 = host::make_mod<uint32_t>(q < (1ull << 32) ? q : 3, (int)I.logn);
    p->mod64 = host::make_mod<uint64_t>(q, (int)I.logn);
    I.barrett_k = p->mod64.k;
    I.barrett_mu = p->mod64.mu;
    p->one_tw = host::make_tw<uint64_t>(1, q);
    p->ninv_tw = host::make_tw<uint64_t>(I.n_inv, q);
    if (root_is_psi) {
        p->psi_pow = host::powers(I.psi, n, q);
        p->psi_inv_pow = host::powers(I.psi_inv, n, q);
    }
    int rc = I.word_bytes == 4 ? build_tables<uint32_t>(p) : build_tables<uint64_t>(p);
    if (rc != TNTT_OK) { tntt_plan_destroy(p); return rc; }

    // kernels for this (word, n)
    if (I.omega_is_primitive) {
        int c = 0;
        const TransformVariant *tv = transform_variants(&c);
        for (int i = 0; i < c; ++i)
            if (tv[i].word_bytes == I.word_bytes && tv[i].logn == (int)I.logn && tv[i].red == I.lazy_reduce) {
                if (tv[i].red && !host::lazy_pass_ok<uint64_t>(q, tv[i].logr)) continue;
                p->xform = &tv[i];
                cudaError_t e = tv[i].prepare();
                if (e != cudaSuccess) { tntt_plan_destroy(p); return fail(TNTT_CUDA_ERROR, "prepare %s: %s", tv[i].name, cudaGetErrorString(e)); }
                break;
            }
    }
    if (I.omega_is_primitive) {
        int c = 0;
        const SpectrumVariant *sv = spectrum_variants(&c);
        for (int i = 0; i < c; ++i)
            if (sv[i].word_bytes == I.word_bytes && sv[i].logn == (int)I.logn && p->cyc_ct_last[sv[i].logr] &&
                (sv[i].red == 2 ? (I.lazy_reduce && I.solinas) : sv[i].red == I.lazy_reduce)) {
                if (sv[i].red && !host::lazy_pass_ok<uint64_t>(q, sv[i].logr)) continue;
                cudaError_t e = sv[i].prepare();
                if (e != cudaSuccess) { tntt_plan_destroy(p); return fail(TNTT_CUDA_ERROR, "prepare %s: %s", sv[i].name, cudaGetErrorString(e)); }
                p->spectrum = &sv[i];
                break;
            }
    }
    I.spectrum = (p->spectrum && I.has_psi && p->fwd_last[p->spectrum->logr]) ? 1 : 0;
    const std::vector<PolymulVariant> &vs = all_variants();
    for (size_t i = 0; i < vs.size(); ++i)
        if (tntt_variant_matches(p, (int)i)) {
            cudaError_t e = vs[i].prepare();
            if (e != cudaSuccess) { tntt_plan_destroy(p); return fail(TNTT_CUDA_ERROR, "prepare %s: %s", vs[i].name, cudaGetErrorString(e)); }
        }
    I.default_variant = choose_default_variant(p);
    I.fused = I.default_variant >= 0 ? 1 : 0;
    // rows on thread-block clusters (N = 16384, 32768): the transform-domain kernels have the cluster shape of the fused
    // kernel, which choose_default_variant() only accepts when the device can co-schedule it (cudaOccupancyMaxActiveClusters)
    if (p->spectrum && p->spectrum->cluster && !(I.default_variant >= 0 && vs[I.default_variant].cluster)) {
        p->spectrum = nullptr;
        I.spectrum = 0;
    }
    choose_small_batch_variants(p);
    CUDA_TRY(cudaDeviceSynchronize());
    *out = p;
    return TNTT_OK;
}

int check_io(const tntt_plan *p, const void *a, const void *b, size_t batch = 1) {
    if (!p) return fail(TNTT_BAD_ARG, "plan is null");
    if (batch == 0) return TNTT_OK;   // empty batches may come with null pointers
    if (!a || !b) return fail(TNTT_BAD_ARG, "null data pointer");
    if (((uintptr_t)a | (uintptr_t)b) & 15) return fail(TNTT_BAD_ARG, "data pointers must be 16-byte aligned");
    return TNTT_OK;
}

// literal constant-geometry transform: bit reversal + log n stages through two scratch buffers
int generic_transform(const tntt_plan *p, const void *in, void *out, size_t batch, bool inverse, int flags, cudaStream_t st) {
    const int wb = p->info.word_bytes, logn = (int)p->info.logn;
    const size_t bytes = batch * p->info.n * (size_t)wb;
    void *s0 = nullptr, *s1 = nullptr;
    struct Scratch {   // released on every path out of this function, error returns included
        void *&a, *&b;
        cudaStream_t st;
        ~Scratch() { if (a) cudaFreeAsync(a, st); if (b) cudaFreeAsync(b, st); }
    } scratch{s0, s1, st};
    CUDA_TRY(cudaMallocAsync(&s0, bytes, st));
    CUDA_TRY(cudaMallocAsync(&s1, bytes, st));
    const void *src = in;
    if (flags & TNTT_REDUCE_INPUT) { CUDA_TRY(launch_reduce(wb, src, s1, batch * p->info.n, p->mod(), st)); src = s1; }
    if (!inverse && (flags & TNTT_TWIST)) { CUDA_TRY(launch_mul_table(wb, src, s1, batch, logn, p->pre_twist, p->mod(), st)); src = s1; }
    CUDA_TRY(launch_bit_reverse(wb, src, s0, batch, logn, st));
    void *cur = s0, *nxt = s1;
    for (int stage = 1; stage <= logn; ++stage) {
        CUDA_TRY(launch_cg_stage(wb, cur, nxt, batch, logn, stage, inverse ? p->omega_inv_pow : p->omega_pow, p->mod(), st));
        void *t = cur; cur = nxt; nxt = t;
    }
    if (inverse && (flags & TNTT_TWIST)) CUDA_TRY(launch_mul_table(wb, cur, out, batch, logn, p->post_untwist, p->mod(), st));
    else if (inverse) CUDA_TRY(launch_scale(wb, cur, out, batch * p->info.n, p->ninv_tw.w, wb == 4 ? host::make_tw<uint32_t>(p->info.n_inv, p->info.q).wp : p->ninv_tw.wp, p->mod(), st));
    else CUDA_TRY(cudaMemcpyAsync(out, cur, bytes, cudaMemcpyDeviceToDevice, st));
    return TNTT_OK;
}

template <typename W>
int fast_transform(const tntt_plan *p, const void *in, void *out, size_t batch, bool inverse, int flags, cudaStream_t st) {
    TransformTables<W> tt;
    tt.dit.pyr = (const Tw<W> *)(inverse ? p->inv_pyr : p->cyc_fwd_pyr);
    if constexpr (sizeof(W) == 4) memcpy(tt.dit.head, p->head32[inverse ? 1 : 2], sizeof tt.dit.head);
    else memcpy(tt.dit.head, p->head64[inverse ? 1 : 2], sizeof tt.dit.head);
    tt.pre = (!inverse && (flags & TNTT_TWIST)) ? (const Tw<W> *)p->pre_twist : nullptr;
    tt.post = (inverse && (flags & TNTT_TWIST)) ? (const Tw<W> *)p->post_untwist : nullptr;
    tt.post_uniform = host::make_tw<W>(inverse ? p->info.n_inv : 1, p->info.q);
    tt.reduce_input = (flags & TNTT_REDUCE_INPUT) ? 1 : 0;
    CUDA_TRY(p->xform->launch(in, out, batch, &tt, p->mod(), st));
    return TNTT_OK;
}

// natural-order transforms on the fused kernel's passes (canonical inputs): forward = Cooley-Tukey high bit first +
// permuted store, inverse = permuted load + decimation in time; twist = merged-psi tables / psi^-i N^-1 store table
template <typename W>
int natural_transform(const tntt_plan *p, const void *in, void *out, size_t batch, bool inverse, int flags, cudaStream_t st) {
    const SpectrumVariant &v = *p->spectrum;
    const bool twist = (flags & TNTT_TWIST) != 0;
    PolymulTables<W> tb{};
    tb.inv.pyr = (const Tw<W> *)p->inv_pyr;
    const int hf = twist ? 0 : 3;
    tb.fwd_pyr = (const Tw<W> *)(twist ? p->fwd_pyr : p->cyc_ct_pyr);
    tb.fwd_last = (const Tw<W> *)(twist ? p->fwd_last[v.logr] : p->cyc_ct_last[v.logr]);
    if constexpr (sizeof(W) == 4) { memcpy(tb.fwd_head, p->head32[hf], sizeof tb.fwd_head); memcpy(tb.inv.head, p->head32[1], sizeof tb.inv.head); }
    else { memcpy(tb.fwd_head, p->head64[hf], sizeof tb.fwd_head); memcpy(tb.inv.head, p->head64[1], sizeof tb.inv.head); }
    cudaError_t e;
    if (!inverse) e = v.forward_natural(in, out, batch, &tb, p->mod(), st);
    else {
        const Tw<W> ninv = host::make_tw<W>(p->info.n_inv, p->info.q);
        e = v.inverse_natural(in, out, batch, &tb, twist ? p->post_untwist : nullptr, ninv.w, ninv.wp, p->mod(), st);
    }
    if (e != cudaSuccess) return fail(TNTT_CUDA_ERROR, "%s: %s", v.name, cudaGetErrorString(e));
    return TNTT_OK;
}

int transform(const tntt_plan *p, const void *in, void *out, size_t batch, bool inverse, int flags, void *stream) {
    int rc = check_io(p, in, out, batch);
    if (rc) return rc;
    if (flags & ~(TNTT_TWIST | TNTT_REDUCE_INPUT)) return fail(TNTT_BAD_ARG, "unknown flag bits 0x%x", flags);
    if ((flags & TNTT_TWIST) && !p->info.has_psi) return fail(TNTT_BAD_ARG, "TNTT_TWIST needs a plan created from psi");
    if (batch == 0) return TNTT_OK;
    DeviceSetter ds(p->info.device);
    cudaStream_t st = (cudaStream_t)stream;
    if (p->spectrum && p->spectrum->forward_natural && !(flags & TNTT_REDUCE_INPUT) && (!(flags & TNTT_TWIST) || p->info.spectrum))
        return p->info.word_bytes == 4 ? natural_transform<uint32_t>(p, in, out, batch, inverse, flags, st)
                                       : natural_transform<uint64_t>(p, in, out, batch, inverse, flags, st);
    if (p->xform) return p->info.word_bytes == 4 ? fast_transform<uint32_t>(p, in, out, batch, inverse, flags, st)
                                                : fast_transform<uint64_t>(p, in, out, batch, inverse, flags, st);
    return generic_transform(p, in, out, batch, inverse, flags, st);
}

template <typename W> int launch_variant(const tntt_plan *p, const PolymulVariant &v, const void *a, const void *b, void *c,
                                         size_t batch, cudaStream_t st) {
    PolymulTables<W> tb;
    tb.fwd_pyr = (const Tw<W> *)p->fwd_pyr;
    tb.fwd_last = (const Tw<W> *)p->fwd_last[v.logr];
    // the Montgomery pointwise product of red 0/1 leaves a factor 2^-BITS for the store table to undo; the Solinas one does not
    tb.post = (const Tw<W> *)(v.red >= 2 ? p->post_untwist : p->post_mont);
    tb.inv.pyr = (const Tw<W> *)p->inv_pyr;
    if constexpr (sizeof(W) == 4) { memcpy(tb.fwd_head, p->head32[0], sizeof tb.fwd_head); memcpy(tb.inv.head, p->head32[1], sizeof tb.inv.head); }
    else { memcpy(tb.fwd_head, p->head64[0], sizeof tb.fwd_head); memcpy(tb.inv.head, p->head64[1], sizeof tb.inv.head); }
    CUDA_TRY(v.launch(a, b, c, batch, &tb, p->mod(), st));
    return TNTT_OK;
}

template <typename W> void fill_polymul_tables(const tntt_plan *p, int logr, PolymulTables<W> &tb) {
    tb.fwd_pyr = (const Tw<W> *)p->fwd_pyr;
    tb.fwd_last = (const Tw<W> *)p->fwd_last[logr];
    tb.post = (const Tw<W> *)p->post_mont;
    tb.inv.pyr = (const Tw<W> *)p->inv_pyr;
    if constexpr (sizeof(W) == 4) { memcpy(tb.fwd_head, p->head32[0], sizeof tb.fwd_head); memcpy(tb.inv.head, p->head32[1], sizeof tb.inv.head); }
    else { memcpy(tb.fwd_head, p->head64[0], sizeof tb.fwd_head); memcpy(tb.inv.head, p->head64[1], sizeof tb.inv.head); }
}
// op: 0 = forward, 1 = inverse, 2 = polymul with b in the transform domain
template <typename W> int spectrum_op(const tntt_plan *p, int op, const void *a, const void *b, void *out, size_t batch,
                                      size_t b_stride, cudaStream_t st) {
    PolymulTables<W> tb;
    fill_polymul_tables<W>(p, p->spectrum->logr, tb);
    if (p->spectrum->red == 2) tb.post = (const Tw<W> *)p->post_untwist;   // Solinas pointwise product: no 2^-BITS to undo
    cudaError_t e = cudaSuccess;
    if (op == 0) e = p->spectrum->forward(a, out, batch, &tb, p->mod(), st);
    else if (op == 1) e = p->spectrum->inverse(a, out, batch, &tb, p->post_untwist, p->mod(), st);
    else e = p->spectrum->polymul(a, b, out, batch, b_stride, &tb, p->mod(), st);
    if (e != cudaSuccess) return fail(TNTT_CUDA_ERROR, "%s: %s", p->spectrum->name, cudaGetErrorString(e));
    return TNTT_OK;
}
int spectrum_entry(const tntt_plan *p, int op, const void *a, const void *b, void *out, size_t batch, size_t b_rows, void *stream) {
    int rc = check_io(p, a, op == 2 ? b : a, batch);
    if (rc) return rc;
    if (!p->info.spectrum) return fail(TNTT_UNSUPPORTED_N, "no transform-domain kernels for this plan (needs psi and a fused size: n in {256, 512, ..., 32768})");
    if (batch == 0) return TNTT_OK;
    if (!out || ((uintptr_t)out & 15)) return fail(TNTT_BAD_ARG, "output must be a 16-byte aligned device pointer");
    if (op == 2 && b_rows != 1 && b_rows != batch) return fail(TNTT_BAD_ARG, "b_rows must be 1 (shared spectrum) or the batch size");
    DeviceSetter ds(p->info.device);
    const size_t b_stride = (op == 2 && b_rows == batch) ? p->info.n : 0;   // 0: every row reads the one shared spectrum
    return p->info.word_bytes == 4 ? spectrum_op<uint32_t>(p, op, a, b, out, batch, b_stride, (cudaStream_t)stream)
                                   : spectrum_op<uint64_t>(p, op, a, b, out, batch, b_stride, (cudaStream_t)stream);
}

int generic_polymul(const tntt_plan *p, const void *a, const void *b, void *c, size_t batch, cudaStream_t st) {
    const int wb = p->info.word_bytes;
    const size_t bytes = batch * p->info.n * (size_t)wb;
    void *fa = nullptr, *fb = nullptr;
    CUDA_TRY(cudaMallocAsync(&fa, bytes, st));
    CUDA_TRY(cudaMallocAsync(&fb, bytes, st));
    int rc = generic_transform(p, a, fa, batch, false, TNTT_TWIST, st);
    if (!rc) rc = generic_transform(p, b, fb, batch, false, TNTT_TWIST, st);
    if (!rc && launch_pointwise(wb, fa, fb, fa, batch * p->info.n, p->mod(), st) != cudaSuccess) rc = fail(TNTT_CUDA_ERROR, "pointwise launch failed");
    if (!rc) rc = generic_transform(p, fa, c, batch, true, TNTT_TWIST, st);
    cudaFreeAsync(fa, st);
    cudaFreeAsync(fb, st);
    return rc;
}

}  // namespace

extern "C" {
#pragma GCC visibility push(default)

int tntt_version(void) { return TNTT_VERSION; }
size_t tntt_plan_info_size(void) { return sizeof(tntt_plan_info); }
const char *tntt_last_error(void) { return g_err.c_str(); }

int tntt_plan_create(tntt_plan **out, int device, uint32_t n, uint64_t q, uint64_t root, int root_is_psi) {
    return create_plan(out, device, n, q, root, root_is_psi);
}

int tntt_plan_create_from_hex(tntt_plan **out, int device, uint32_t n, uint64_t q, const char *fwd_hex_path,
                              const char *inv_hex_path) {
    if (!out || !fwd_hex_path) return fail(TNTT_BAD_ARG, "null argument");
    auto load = [&](const char *path, std::vector<uint64_t> &v) -> int {
        FILE *f = fopen(path, "r");
        if (!f) return fail(TNTT_IO_ERROR, "cannot open %s", path);
        char line[128];
        while (fgets(line, sizeof line, f)) {
            char *end = nullptr;
            if (line[0] == '\n' || line[0] == '\r' || line[0] == 0) continue;
            const unsigned long long val = strtoull(line, &end, 16);
            if (end == line) { fclose(f); return fail(TNTT_IO_ERROR, "%s: bad hex line '%s'", path, line); }
            v.push_back(val);
        }
        fclose(f);
        if (v.size() != n) return fail(TNTT_IO_ERROR, "%s: %zu entries, expected %u", path, v.size(), n);
        return TNTT_OK;
    };
    std::vector<uint64_t> fwd, inv;
    int rc = load(fwd_hex_path, fwd);
    if (rc) return rc;
    if (n < 2 || fwd[0] != 1) return fail(TNTT_BAD_ROOT, "forward table must start with psi^0 = 1");
    const uint64_t psi = fwd[1];
    if (psi >= q) return fail(TNTT_BAD_ROOT, "table entry exceeds q");
    const std::vector<uint64_t> want = host::powers(psi, n, q);
    for (uint32_t k = 0; k < n; ++k)
        if (fwd[k] != want[k]) return fail(TNTT_BAD_ROOT, "%s: entry %u is not psi^%u", fwd_hex_path, k, k);
    if (inv_hex_path) {
        rc = load(inv_hex_path, inv);
        if (rc) return rc;
        const std::vector<uint64_t> wi = host::powers(host::modinv(psi, q), n, q);
        for (uint32_t k = 0; k < n; ++k)
            if (inv[k] != wi[k]) return fail(TNTT_BAD_ROOT, "%s: entry %u is not psi^-%u", inv_hex_path, k, k);
    }
    return create_plan(out, device, n, q, psi, 1);
}

int tntt_plan_write_hex(const tntt_plan *plan, const char *path, int inverse, int hex_digits, int uppercase) {
    if (!plan || !path) return fail(TNTT_BAD_ARG, "null argument");
    if (!plan->info.has_psi) return fail(TNTT_BAD_ARG, "plan has no psi");
    if (hex_digits < 1 || hex_digits > 16) return fail(TNTT_BAD_ARG, "hex_digits out of range");
    FILE *f = fopen(path, "w");
    if (!f) return fail(TNTT_IO_ERROR, "cannot open %s for writing", path);
    for (uint64_t v : (inverse ? plan->psi_inv_pow : plan->psi_pow)) fprintf(f, uppercase ? "%0*llX\n" : "%0*llx\n", hex_digits, (unsigned long long)v);
    fclose(f);
    return TNTT_OK;
}

int tntt_plan_info_get(const tntt_plan *plan, tntt_plan_info *info) {
    if (!plan || !info) return fail(TNTT_BAD_ARG, "null argument");
    *info = plan->info;
    return TNTT_OK;
}

int tntt_plan_destroy(tntt_plan *p) {
    if (!p) return TNTT_OK;
    DeviceSetter ds(p->info.device);
    void *bufs[] = {p->fwd_pyr, p->inv_pyr, p->post_mont, p->cyc_fwd_pyr, p->pre_twist, p->post_untwist, p->omega_pow, p->omega_inv_pow};
    for (void *b : bufs) if (b) cudaFree(b);
    for (void *b : p->fwd_last) if (b) cudaFree(b);
    if (p->cyc_ct_pyr) cudaFree(p->cyc_ct_pyr);
    for (void *t : p->cyc_ct_last) if (t) cudaFree(t);
    for (int s = 0; s < tntt_plan::kSlots; ++s) {
        for (void *b : p->pipe_buf[s]) if (b) cudaFree(b);
        if (p->pipe_stream[s]) cudaStreamDestroy(p->pipe_stream[s]);
    }
    delete p;
    return TNTT_OK;
}

int tntt_forward(const tntt_plan *plan, const void *in, void *out, size_t batch, int flags, void *stream) {
    return transform(plan, in, out, batch, false, flags, stream);
}
int tntt_inverse(const tntt_plan *plan, const void *in, void *out, size_t batch, int flags, void *stream) {
    return transform(plan, in, out, batch, true, flags, stream);
}

int tntt_pointwise(const tntt_plan *p, const void *a, const void *b, void *c, size_t batch, void *stream) {
    int rc = check_io(p, a, b, batch);
    if (rc) return rc;
    if (batch == 0) return TNTT_OK;
    if (!c) return fail(TNTT_BAD_ARG, "null data pointer");
    DeviceSetter ds(p->info.device);
    CUDA_TRY(launch_pointwise(p->info.word_bytes, a, b, c, batch * p->info.n, p->mod(), (cudaStream_t)stream));
    return TNTT_OK;
}

int tntt_spectrum_forward(const tntt_plan *p, const void *in, void *out, size_t batch, void *stream) {
    return spectrum_entry(p, 0, in, nullptr, out, batch, 0, stream);
}
int tntt_spectrum_inverse(const tntt_plan *p, const void *in, void *out, size_t batch, void *stream) {
    return spectrum_entry(p, 1, in, nullptr, out, batch, 0, stream);
}
int tntt_polymul_spectrum(const tntt_plan *p, const void *a, const void *b_spectrum, void *c, size_t batch, size_t b_rows,
                          void *stream) {
    return spectrum_entry(p, 2, a, b_spectrum, c, batch, b_rows, stream);
}

int tntt_variant_count(void) { return (int)all_variants().size(); }
int tntt_variant_describe(int variant, char *buf, size_t buflen) {
    if (variant < 0 || variant >= tntt_variant_count() || !buf) return fail(TNTT_BAD_ARG, "bad variant");
    const PolymulVariant &v = all_variants()[variant];
    cudaFuncAttributes attr{};
    int occ = 0;
    const bool have = v.attributes(&attr, &occ) == cudaSuccess;
    if (!have) cudaGetLastError();
    snprintf(buf, buflen, "%s word=%d n=%d r=%d ppc=%d na=%d red=%d threads=%d smem=%zu regs=%d local=%zu ctas_per_sm=%d",
             v.name, v.word_bytes, 1 << v.logn, 1 << v.logr, v.ppc, v.na, v.red, v.threads, v.smem,
             have ? attr.numRegs : -1, have ? attr.localSizeBytes : (size_t)0, have ? occ : -1);
    return TNTT_OK;
}
int tntt_variant_matches(const tntt_plan *p, int variant) {
    if (!p || variant < 0 || variant >= tntt_variant_count()) return 0;
    const PolymulVariant &v = all_variants()[variant];
    if (!p->info.has_psi || p->info.literal_only || v.word_bytes != p->info.word_bytes || v.logn != (int)p->info.logn) return 0;
    // red 2 = the Solinas-form reductions: an alternative to red 1 for the one modulus they are written for
    // red 3 = Barrett products of canonical values: serves every modulus of the 64-bit paths
    if (v.red == 3) return v.word_bytes == 8 ? 1 : 0;
    if (v.red == 2 ? !(p->info.lazy_reduce && p->info.solinas) : v.red != p->info.lazy_reduce) return 0;
    if (v.red && !host::lazy_pass_ok<uint64_t>(p->info.q, v.logr)) return 0;
    return 1;
}
int tntt_plan_set_default_variant(tntt_plan *p, int variant) {
    if (!tntt_variant_matches(p, variant)) return fail(TNTT_BAD_ARG, "variant %d does not match the plan", variant);
    p->info.default_variant = variant;
    p->info.cluster_variant = p->info.small_variant = -1;   // an explicit choice switches the batch-size dispatch off
    p->info.cluster_batch_max = p->info.small_batch_max = 0;
    return TNTT_OK;
}

int tntt_polymul_variant(const tntt_plan *p, int variant, const void *a, const void *b, void *c, size_t batch, void *stream) {
    int rc = check_io(p, a, b, batch);
    if (rc) return rc;
    if (batch == 0) return TNTT_OK;
    if (!c || ((uintptr_t)c & 15)) return fail(TNTT_BAD_ARG, "c must be a 16-byte aligned device pointer");
    if (!tntt_variant_matches(p, variant)) return fail(TNTT_BAD_ARG, "variant %d does not match the plan", variant);
    DeviceSetter ds(p->info.device);
    const PolymulVariant &v = all_variants()[variant];
    return p->info.word_bytes == 4 ? launch_variant<uint32_t>(p, v, a, b, c, batch, (cudaStream_t)stream)
                                   : launch_variant<uint64_t>(p, v, a, b, c, batch, (cudaStream_t)stream);
}

int tntt_polymul(const tntt_plan *p, const void *a, const void *b, void *c, size_t batch, void *stream) {
    int rc = check_io(p, a, b, batch);
    if (rc) return rc;
    if (!p->info.has_psi) return fail(TNTT_BAD_ARG, "polymul needs a plan created from psi");
    if (batch == 0) return TNTT_OK;
    if (!c || ((uintptr_t)c & 15)) return fail(TNTT_BAD_ARG, "c must be a 16-byte aligned device pointer");
    if (p->info.default_variant >= 0) {
        int v = p->info.default_variant;
        if (p->info.cluster_variant >= 0 && batch <= (size_t)p->info.cluster_batch_max) v = p->info.cluster_variant;
        else if (p->info.small_variant >= 0 && batch <= (size_t)p->info.small_batch_max) v = p->info.small_variant;
        return tntt_polymul_variant(p, v, a, b, c, batch, stream);
    }
    DeviceSetter ds(p->info.device);
    return generic_polymul(p, a, b, c, batch, (cudaStream_t)stream);
}

// The host pipeline behind tntt_polymul_host (b_spectrum == nullptr: a and b both come from the host) and
// tntt_polymul_spectrum_host (b is a spectrum that already lives on the device; b_rows = batch or 1).
static int polymul_host_pipeline(tntt_plan *p, const void *a, const void *b, const void *b_spectrum, size_t b_rows, void *c,
                                 size_t batch) {
    DeviceSetter ds(p->info.device);
    std::lock_guard<std::mutex> lock(p->pipe_mu);
    const size_t row_bytes = (size_t)p->info.n * p->info.word_bytes;
    // Chunks ramp up from 1/16 of the full chunk, stay at the full chunk, and ramp down again: the first kernel
    // starts after a short copy and the last kernel + copy-back is short, so that a blocking call is PCIe-bound
    // for all but ~0.2 ms (a fixed chunk size pays one whole chunk of H2D before, and one of D2H after, the
    // overlapped part).  Full chunk: 64 MiB per operand, measured best on PCIe gen5 (TNTT_HOST_CHUNK_MB overrides).
    size_t chunk_mb = 64;
    if (const char *env = getenv("TNTT_HOST_CHUNK_MB")) { const long v = atol(env); if (v >= 1 && v <= 1024) chunk_mb = (size_t)v; }
    size_t full = (chunk_mb << 20) / row_bytes;
    if (full < 1) full = 1;
    if (full > batch) full = batch;
    if (p->pipe_rows < full) {
        for (int s = 0; s < tntt_plan::kSlots; ++s) {
            if (!p->pipe_stream[s]) CUDA_TRY(cudaStreamCreateWithFlags(&p->pipe_stream[s], cudaStreamNonBlocking));
            for (int k = 0; k < 3; ++k) {
                if (p->pipe_buf[s][k]) CUDA_TRY(cudaFree(p->pipe_buf[s][k]));
                p->pipe_buf[s][k] = nullptr;
                CUDA_TRY(cudaMalloc(&p->pipe_buf[s][k], full * row_bytes));
            }
        }
        p->pipe_rows = full;
    }
    std::vector<size_t> head, tail;
    {
        const size_t first = full / 16 > 0 ? full / 16 : 1;
        size_t h = first, t = first, rem = batch;
        while (rem) {
            size_t n = h < rem ? h : rem;
            head.push_back(n);
            rem -= n;
            h = 2 * h < full ? 2 * h : full;
            if (rem && t < full) {
                n = t < rem ? t : rem;
                tail.push_back(n);
                rem -= n;
                t = 2 * t < full ? 2 * t : full;
            }
        }
        head.insert(head.end(), tail.rbegin(), tail.rend());
    }
    const char *pa = (const char *)a, *pb = (const char *)b;
    char *pc = (char *)c;
    int slot = 0, rc = TNTT_OK;
    size_t r0 = 0;
    cudaError_t e = cudaSuccess;
    const char *what = "";
    for (const size_t nr : head) {
        cudaStream_t st = p->pipe_stream[slot];
        what = "cudaMemcpyAsync (host to device)";
        if ((e = cudaMemcpyAsync(p->pipe_buf[slot][0], pa + r0 * row_bytes, nr * row_bytes, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
        if (b_spectrum) {
            const char *bs = (const char *)b_spectrum + (b_rows == 1 ? 0 : r0 * row_bytes);
            if ((rc = tntt_polymul_spectrum(p, p->pipe_buf[slot][0], bs, p->pipe_buf[slot][2], nr, b_rows == 1 ? 1 : nr, st)) != TNTT_OK) break;
        } else {
            if ((e = cudaMemcpyAsync(p->pipe_buf[slot][1], pb + r0 * row_bytes, nr * row_bytes, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
            if ((rc = tntt_polymul(p, p->pipe_buf[slot][0], p->pipe_buf[slot][1], p->pipe_buf[slot][2], nr, st)) != TNTT_OK) break;
        }
        what = "cudaMemcpyAsync (device to host)";
        if ((e = cudaMemcpyAsync(pc + r0 * row_bytes, p->pipe_buf[slot][2], nr * row_bytes, cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
        r0 += nr;
        slot = (slot + 1) % tntt_plan::kSlots;
    }
    // drain every stream even after a failure: copies already queued still read a/b and write c, and the caller
    // is free to release those buffers as soon as this returns
    for (int s = 0; s < tntt_plan::kSlots; ++s) {
        const cudaError_t se = cudaStreamSynchronize(p->pipe_stream[s]);
        if (se != cudaSuccess && e == cudaSuccess && rc == TNTT_OK) { e = se; what = "cudaStreamSynchronize"; }
    }
    if (rc != TNTT_OK) return rc;
    if (e != cudaSuccess) return fail(TNTT_CUDA_ERROR, "%s: %s", what, cudaGetErrorString(e));
    return TNTT_OK;
}

int tntt_polymul_host(tntt_plan *p, const void *a, const void *b, void *c, size_t batch) {
    if (!p || !a || !b || !c) return fail(TNTT_BAD_ARG, "null argument");
    if (!p->info.has_psi) return fail(TNTT_BAD_ARG, "polymul needs a plan created from psi");
    if (batch == 0) return TNTT_OK;
    return polymul_host_pipeline(p, a, b, nullptr, 0, c, batch);
}

int tntt_polymul_spectrum_host(tntt_plan *p, const void *a, const void *b_spectrum, size_t b_rows, void *c, size_t batch) {
    if (!p || !a || !b_spectrum || !c) return fail(TNTT_BAD_ARG, "null argument");
    if (!p->info.spectrum) return fail(TNTT_BAD_ARG, "this plan has no transform-domain kernels (tntt_plan_info.spectrum == 0)");
    if (batch == 0) return TNTT_OK;
    if (b_rows != 1 && b_rows != batch) return fail(TNTT_BAD_ARG, "b_rows must be 1 or batch");
    if ((uintptr_t)b_spectrum & 15) return fail(TNTT_BAD_ARG, "b_spectrum must be a 16-byte aligned device pointer");
    return polymul_host_pipeline(p, a, nullptr, b_spectrum, b_rows, c, batch);
}

// Single-process multi-GPU form of the same call: plans[i] lives on its own device
// and takes the i-th contiguous range of ceil(batch / nplans) rows; one host thread per device drives that device's
// pipeline; the join is the host barrier.  No data-path collective: rows are independent (cg_ntt.py:78-92).
int tntt_polymul_host_multi(tntt_plan *const *plans, int nplans, const void *a, const void *b, void *c, size_t batch) {
    if (!plans || nplans < 1) return fail(TNTT_BAD_ARG, "need at least one plan");
    for (int i = 0; i < nplans; ++i) {
        if (!plans[i]) return fail(TNTT_BAD_ARG, "plan %d is null", i);
        if (plans[i]->info.n != plans[0]->info.n || plans[i]->info.q != plans[0]->info.q || plans[i]->info.psi != plans[0]->info.psi)
            return fail(TNTT_BAD_ARG, "plan %d is for another ring than plan 0", i);
        for (int j = 0; j < i; ++j)
            if (plans[j] == plans[i] || plans[j]->info.device == plans[i]->info.device)
                return fail(TNTT_BAD_ARG, "plans %d and %d share device %d: one plan per device", j, i, plans[i]->info.device);
    }
    if (batch == 0) return TNTT_OK;
    if (!a || !b || !c) return fail(TNTT_BAD_ARG, "null argument");
    if (nplans == 1) return tntt_polymul_host(plans[0], a, b, c, batch);
    const size_t per = (batch + (size_t)nplans - 1) / (size_t)nplans;
    const size_t row_bytes = (size_t)plans[0]->info.n * plans[0]->info.word_bytes;
    std::vector<int> rcs(nplans, TNTT_OK);
    std::vector<std::string> msgs(nplans);
    std::vector<std::thread> workers;
    for (int i = 0; i < nplans; ++i) {
        const size_t r0 = per * (size_t)i < batch ? per * (size_t)i : batch;
        const size_t rows = batch - r0 < per ? batch - r0 : per;
        if (!rows) continue;
        workers.emplace_back([=, &rcs, &msgs] {
            rcs[i] = tntt_polymul_host(plans[i], (const char *)a + r0 * row_bytes, (const char *)b + r0 * row_bytes,
                                       (char *)c + r0 * row_bytes, rows);
            if (rcs[i] != TNTT_OK) msgs[i] = g_err;     // the message is thread-local: carry it over to the caller
        });
    }
    for (std::thread &t : workers) t.join();
    for (int i = 0; i < nplans; ++i)
        if (rcs[i] != TNTT_OK) return fail(rcs[i], "device %d: %s", plans[i]->info.device, msgs[i].c_str());
    return TNTT_OK;
}

int tntt_cg_stage(const tntt_plan *p, const void *in, void *out, size_t batch, int stage, int inverse, void *stream) {
    int rc = check_io(p, in, out);
    if (rc) return rc;
    if (in == out) return fail(TNTT_BAD_ARG, "cg_stage is out of place");
    if (stage < 1 || stage > (int)p->info.logn) return fail(TNTT_BAD_ARG, "stage %d out of range 1..%u", stage, p->info.logn);
    DeviceSetter ds(p->info.device);
    CUDA_TRY(launch_cg_stage(p->info.word_bytes, in, out, batch, (int)p->info.logn, stage,
                             inverse ? p->omega_inv_pow : p->omega_pow, p->mod(), (cudaStream_t)stream));
    return TNTT_OK;
}
int tntt_bit_reverse(const tntt_plan *p, const void *in, void *out, size_t batch, void *stream) {
    int rc = check_io(p, in, out);
    if (rc) return rc;
    if (in == out) return fail(TNTT_BAD_ARG, "bit_reverse is out of place");
    DeviceSetter ds(p->info.device);
    CUDA_TRY(launch_bit_reverse(p->info.word_bytes, in, out, batch, (int)p->info.logn, (cudaStream_t)stream));
    return TNTT_OK;
}
int tntt_scale(const tntt_plan *p, const void *in, void *out, size_t batch, uint64_t scalar, void *stream) {
    int rc = check_io(p, in, out);
    if (rc) return rc;
    if (scalar >= p->info.q) return fail(TNTT_BAD_ARG, "scalar must be reduced mod q");
    DeviceSetter ds(p->info.device);
    const uint64_t wp = p->info.word_bytes == 4 ? host::make_tw<uint32_t>(scalar, p->info.q).wp
                                                : host::make_tw<uint64_t>(scalar, p->info.q).wp;
    CUDA_TRY(launch_scale(p->info.word_bytes, in, out, batch * p->info.n, scalar, wp, p->mod(), (cudaStream_t)stream));
    return TNTT_OK;
}
int tntt_reduce(const tntt_plan *p, const void *in, void *out, size_t batch, void *stream) {
    int rc = check_io(p, in, out);
    if (rc) return rc;
    DeviceSetter ds(p->info.device);
    CUDA_TRY(launch_reduce(p->info.word_bytes, in, out, batch * p->info.n, p->mod(), (cudaStream_t)stream));
    return TNTT_OK;
}

int tntt_butterfly_batch(int device, uint64_t q, const uint64_t *a, const uint64_t *b, const uint64_t *w, uint64_t *out_a,
                         uint64_t *out_b, size_t count, void *stream) {
    if (!a || !b || !w || !out_a || !out_b) return fail(TNTT_BAD_ARG, "null data pointer");
    if (q < 3 || !(q & 1) || q >= (1ull << 60)) return fail(TNTT_UNSUPPORTED_Q, "q=%llu must be odd, >= 3 and < 2^60", (unsigned long long)q);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(TNTT_NO_DEVICE, "no CUDA device"); }
    DeviceSetter ds(device);
    CUDA_TRY(launch_butterfly(a, b, w, out_a, out_b, count, host::make_mod<uint64_t>(q), (cudaStream_t)stream));
    return TNTT_OK;
}

int tntt_microbench(int device, int kind, double *ops_per_second) {
    if (!ops_per_second) return fail(TNTT_BAD_ARG, "null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(TNTT_NO_DEVICE, "no CUDA device"); }
    DeviceSetter ds(device);
    CUDA_TRY(run_microbench(kind, ops_per_second));
    return TNTT_OK;
}

#pragma GCC visibility pop
}  // extern "C"
