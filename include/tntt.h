/*
 * tntt.h -- C ABI of libtntt.so, the B200 (sm_100a) batched negacyclic-polymul engine.
 *
 * The reference (orhosko/tiny-ntt) has no FFI layer: its boundary for this path is the Python
 * function API of new_reference/cg_ntt.py (+ cg_ntt_8butterfly.py).  tiny-ntt_b200/cg_ntt.py keeps
 * that API and binds the entry points below through ctypes; INTEGRATION.md shows the stub.  Each
 * entry point names the reference interface it replaces (paths relative to the reference root).
 *
 * Conventions
 *  - All `in/out/a/b/c` pointers of the device entry points are DEVICE pointers owned by the
 *    caller: row-major [batch, n] words, natural coefficient order, 16-byte aligned.  A word is
 *    uint32_t when tntt_plan_info().word_bytes == 4 and uint64_t when it is 8.
 *  - Coefficients must be canonical (0 <= x < q) unless TNTT_REDUCE_INPUT is given; outputs are
 *    always canonical.  in == out (in place) is allowed.
 *  - Calls enqueue work on `cuda_stream` (a cudaStream_t; NULL = default stream) and return without
 *    synchronising.  The library never allocates per call on the fused paths (tables live in the
 *    plan) and never frees caller memory.
 *  - A plan's tables are immutable after creation and the device entry points may be called on it from
 *    several host threads / streams.  Two things are NOT thread-safe: tntt_plan_set_default_variant (a
 *    benchmarking knob that rewrites the plan's dispatch fields; call it before sharing the plan) and
 *    tntt_polymul_host / tntt_polymul_spectrum_host, which serialise callers on the plan's one set of staging buffers.
 *    One plan per device.  There is NO CPU fallback: without a CUDA device tntt_plan_create fails
 *    with TNTT_NO_DEVICE.
 *  - Return value: 0 = TNTT_OK, negative = error; tntt_last_error() gives a thread-local message.
 */
#ifndef TNTT_H
#define TNTT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TNTT_VERSION 200 /* 0.2.0 */

enum tntt_status {
    TNTT_OK = 0,
    TNTT_BAD_ARG = -1,        /* null pointer, bad flag, misaligned buffer */
    TNTT_BAD_ROOT = -2,       /* root not reduced mod q (tntt_find_psi: no root exists) */
    TNTT_UNSUPPORTED_N = -3,  /* n is not a power of two in [2, 65536] */
    TNTT_CUDA_ERROR = -4,
    TNTT_UNSUPPORTED_Q = -5,  /* q < 2 or q >= 2^60 (RNS plans: not an odd prime) */
    TNTT_IO_ERROR = -6,       /* hex table file unreadable or malformed */
    TNTT_NO_DEVICE = -7       /* no CUDA device: there is no CPU path */
};

enum tntt_flags {
    TNTT_TWIST = 1,        /* forward: multiply by psi^i first; inverse: multiply by psi^-i last
                              (new_reference/cg_ntt.py:82-83,91-92).  Off = cg_ntt / cg_intt semantics. */
    TNTT_REDUCE_INPUT = 2  /* inputs are arbitrary words: reduce mod q on load (cg_ntt.py:57-59 accepts them) */
};

typedef struct tntt_plan tntt_plan;

typedef struct tntt_plan_info {
    uint32_t n, logn;
    uint64_t q;
    uint64_t psi, psi_inv;     /* 0 when the plan was created from omega only */
    uint64_t omega, omega_inv; /* omega = psi^2 */
    uint64_t n_inv;
    int word_bytes;            /* 4: uint32 coefficients, 8: uint64 */
    int barrett_k;             /* scripts/precompute_constants.py:30-55 */
    uint64_t barrett_mu;
    int has_psi;               /* negacyclic entry points available */
    int omega_is_primitive;    /* 0: only the literal constant-geometry stage path is used */
    int fused;                 /* a fused single-kernel polymul exists for this (n, word) */
    int lazy_reduce;           /* 1: kernels reduce lazily before every pass (q close to 2^60) */
    int default_variant;       /* index into tntt_variant_* used by tntt_polymul */
    int device;
    /* batch-size dispatch of tntt_polymul (-1 = none): batch <= cluster_batch_max -> cluster_variant (one row
     * per thread-block cluster, exchanges through distributed shared memory), else batch <= small_batch_max ->
     * small_variant (one CTA per SM, all registers), else default_variant.  tntt_plan_set_default_variant
     * switches the dispatch off. */
    int cluster_variant, cluster_batch_max;
    int small_variant, small_batch_max;
    int spectrum;              /* 1: tntt_spectrum_forward / _inverse / tntt_polymul_spectrum are available */
    int literal_only;          /* 1: q is not an odd prime, or the root is not primitive (psi^n != -1, omega = 0): the
                                * plan runs the reference's literal stage schedule only (cg_ntt.py:29-92 accepts any
                                * integers; so does this), no fused / transform-domain kernels */
    int solinas;               /* 1: q = 2^60 - 2^14 + 1 (rtl/ntt_poly_mult.sv:16-26): the fused kernels may use its
                                * shift-and-add reductions (variants "red2") instead of the generic lazy ones */
} tntt_plan_info;

/* Ring parameters N, Q of new_reference/cg_ntt.py:5-6 plus the root the caller passes to
 * cg_ntt (omega, root_is_psi = 0) or nwc_poly_mult (psi, root_is_psi = 1).  Computes what
 * scripts/precompute_constants.py (k, mu) and scripts/generate_twiddles.py /
 * generate_inverse_twiddles.py (psi^k tables) compute offline, and uploads the tables. */
int tntt_plan_create(tntt_plan **out, int device, uint32_t n, uint64_t q, uint64_t root, int root_is_psi);

/* Same, from the reference's $readmemh tables (rtl/twiddle_forward*.hex, rtl/twiddle_inverse*.hex):
 * psi = forward[1]; verifies forward[k] == psi^k and inverse[k] == psi^-k for all k < n. */
int tntt_plan_create_from_hex(tntt_plan **out, int device, uint32_t n, uint64_t q, const char *fwd_hex_path,
                              const char *inv_hex_path);

/* Writes the plan's psi^k (inverse = 0) or psi^-k table in the reference's hex format
 * (scripts/generate_twiddles.py:59-77): `hex_digits` digits per line; the shipped N=256 and 60-bit
 * tables are upper case, the 24-bit N=1024/4096 ones lower case. */
int tntt_plan_write_hex(const tntt_plan *plan, const char *path, int inverse, int hex_digits, int uppercase);

int tntt_plan_info_get(const tntt_plan *plan, tntt_plan_info *info);
int tntt_plan_destroy(tntt_plan *plan);

/* cg_ntt(a, omega, q) (new_reference/cg_ntt.py:29-65; rtl/ntt_forward.sv): natural-order cyclic NTT
 * of every row.  With TNTT_TWIST: ntt(twist(a)), i.e. software_benchmark/benchmark_ntt.cpp:207-211. */
int tntt_forward(const tntt_plan *plan, const void *in, void *out, size_t batch, int flags, void *cuda_stream);

/* cg_intt(A, omega, q) (cg_ntt.py:68-75; rtl/ntt_inverse.sv): inverse transform including the N^-1
 * scaling; with TNTT_TWIST also the psi^-i untwist. */
int tntt_inverse(const tntt_plan *plan, const void *in, void *out, size_t batch, int flags, void *cuda_stream);

/* c[i] = a[i] * b[i] mod q with the reference's Barrett reduction (rtl/ntt_pointwise_mult.v:17-42,
 * rtl/barrett_reduction.v:23-29; cg_ntt.py:88). */
int tntt_pointwise(const tntt_plan *plan, const void *a, const void *b, void *c, size_t batch, void *cuda_stream);

/* nwc_poly_mult(a, b, psi) (cg_ntt.py:78-92; rtl/ntt_poly_mult.sv;
 * software_benchmark/benchmark_ntt.cpp:194-205): c = a*b in Z_q[x]/(x^n+1), one fused kernel,
 * one HBM round trip per polynomial. */
int tntt_polymul(const tntt_plan *plan, const void *a, const void *b, void *c, size_t batch, void *cuda_stream);

/* Same through HOST buffers (pinned for full overlap): chunks ramping up to 64 MiB per operand
 * (TNTT_HOST_CHUNK_MB overrides), H2D -> kernel -> D2H rotating over four streams / buffer sets.  The GPU
 * analogue of the RoCC load/start/read command sequence (chipyard/ntt-test.c:110-169).  Blocks until c is
 * complete; on an error every queued copy is drained before the call returns. */
int tntt_polymul_host(tntt_plan *plan, const void *a_host, const void *b_host, void *c_host, size_t batch);
/* The same over several GPUs from ONE process (per-device plan + streams, batch split
 * into contiguous ranges of ceil(batch / nplans) rows, host barrier): plans[i] must be plans of the same ring on
 * distinct devices.  One host thread per device runs that device's pipeline; returns when all of c is complete. */
int tntt_polymul_host_multi(tntt_plan *const *plans, int nplans, const void *a_host, const void *b_host, void *c_host,
                            size_t batch);

/* Literal constant-geometry schedule, one stage per call (cg_ntt.py:49-59,
 * rtl/ntt_cg_address_gen.v:57-117): out[i] = in[2i] + w*in[2i+1], out[i+n/2] = in[2i] - w*in[2i+1],
 * w = root^((n>>stage) * (i / (n>>stage))), stage = 1..log2(n); root = omega (inverse = 0) or
 * omega^-1.  Used for verbose=True traces and for any n without a fused kernel.  in != out. */
int tntt_cg_stage(const tntt_plan *plan, const void *in, void *out, size_t batch, int stage, int inverse,
                  void *cuda_stream);
/* out[bitrev(i)] = in[i] (cg_ntt.py:21-26; rtl/ntt_coeff_banks.v:43-53,112).  in != out. */
int tntt_bit_reverse(const tntt_plan *plan, const void *in, void *out, size_t batch, void *cuda_stream);
/* out[i] = in[i] * scalar mod q (the N^-1 pass of cg_ntt.py:74-75; rtl/ntt_inverse.sv:375-386). */
int tntt_scale(const tntt_plan *plan, const void *in, void *out, size_t batch, uint64_t scalar, void *cuda_stream);
/* out[i] = in[i] mod q for arbitrary words. */
int tntt_reduce(const tntt_plan *plan, const void *in, void *out, size_t batch, void *cuda_stream);

/* (a + w*b, a - w*b) mod q on `count` independent lanes of canonical uint64 values (device pointers):
 * butterfly / butterfly_batch of new_reference/cg_ntt_8butterfly.py:8-27; rtl/ntt_butterfly.v:43-72. */
int tntt_butterfly_batch(int device, uint64_t q, const uint64_t *a, const uint64_t *b, const uint64_t *w, uint64_t *out_a,
                         uint64_t *out_b, size_t count, void *cuda_stream);

/* Operands kept in the transform domain (one forward transform per operand, reused across many products: the
 * RLWE use the reference targets, reports/final-report.tex:571-610).  A "spectrum" row holds ntt(twist(a))
 * (new_reference/cg_ntt.py:82-87; forward_ntt_bench, software_benchmark/benchmark_ntt.cpp:207-211) as canonical
 * values in a plan-specific order: the order in which the fused kernel holds the transform in registers, so that
 * neither producer nor consumer pays a bit reversal or a twist pass.  Treat the order as opaque; it is a fixed
 * permutation of the natural-order transform, so element-wise work (tntt_pointwise, additions) applies directly.
 *   tntt_spectrum_forward : coefficients -> spectrum
 *   tntt_spectrum_inverse : spectrum -> coefficients, i.e. untwist(cg_intt(.)) (cg_ntt.py:90-92)
 *   tntt_polymul_spectrum : c = a * b in Z_q[x]/(x^n+1) with b given as spectrum; b_rows = batch (one per row) or
 *                           1 (one spectrum shared by the whole batch).  Bit-identical to tntt_polymul.
 * Available when tntt_plan_info.spectrum is 1 (plans created from psi with n in {256, 512, ..., 32768}). */
int tntt_spectrum_forward(const tntt_plan *plan, const void *in, void *out, size_t batch, void *cuda_stream);
int tntt_spectrum_inverse(const tntt_plan *plan, const void *in, void *out, size_t batch, void *cuda_stream);
int tntt_polymul_spectrum(const tntt_plan *plan, const void *a, const void *b_spectrum, void *c, size_t batch,
                          size_t b_rows, void *cuda_stream);
/* tntt_polymul_host with the second operand cached: a and c are HOST buffers, b_spectrum is a DEVICE buffer of
 * b_rows = 1 or batch spectra written by tntt_spectrum_forward (complete before this call: the pipeline runs on the
 * plan's own streams).  The fixed-key pattern of reports/final-report.tex:571-610 from host memory: one third less
 * PCIe traffic than tntt_polymul_host and the two directions carry the same load.  Same pipeline, same blocking and
 * error behaviour; c is bit-identical to tntt_polymul_host(a, b). */
int tntt_polymul_spectrum_host(tntt_plan *plan, const void *a_host, const void *b_spectrum, size_t b_rows, void *c_host,
                               size_t batch);

/* Kernel variants of the fused polymul (tile shape, operands side by side, ...), for benchmarking. */
int tntt_variant_count(void);
int tntt_variant_describe(int variant, char *buf, size_t buflen);   /* "u64 n=4096 r=16 ppc=1 na=1 red=1 ..." */
int tntt_variant_matches(const tntt_plan *plan, int variant);        /* 1 if usable with this plan */
int tntt_polymul_variant(const tntt_plan *plan, int variant, const void *a, const void *b, void *c, size_t batch,
                         void *cuda_stream);
int tntt_plan_set_default_variant(tntt_plan *plan, int variant);

/* ---- multi-modulus (RNS) batches: SURVEY.md section 8 f3; reports/final-report.tex:1811-1817 ----
 * An RNS polynomial batch is [limbs][batch][n] words: limb l holds the residues mod q[l].  All limbs must need
 * the same word size (see tntt_plan_info.word_bytes).  The tables of every limb (the contents of
 * scripts/generate_twiddles.py:29-41 / generate_inverse_twiddles.py:48-61 in kernel order, with their Shoup
 * companions) are generated on the device by one kernel launch from (q[l], psi[l]). */
typedef struct tntt_rns_plan tntt_rns_plan;
int tntt_rns_plan_create(tntt_rns_plan **out, int device, uint32_t n, const uint64_t *q, const uint64_t *psi, int limbs);
void tntt_rns_plan_destroy(tntt_rns_plan *plan);
int tntt_rns_plan_limbs(const tntt_rns_plan *plan);
int tntt_rns_plan_word_bytes(const tntt_rns_plan *plan);
const char *tntt_rns_plan_kernel(const tntt_rns_plan *plan);       /* name of the kernel shape in use */
size_t tntt_rns_plan_table_bytes(const tntt_rns_plan *plan);       /* device memory held by the tables */
/* c[l] = a[l] * b[l] in Z_{q_l}[x]/(x^n+1) for every limb: ONE kernel launch per 16 limbs (the limb index is
 * blockIdx.y; tables and modulus constants are picked out of the kernel parameters). */
int tntt_rns_polymul(const tntt_rns_plan *plan, const void *a, const void *b, void *c, size_t batch, void *cuda_stream);
/* Operands kept in the transform domain, all limbs per launch (the multi-modulus forms of tntt_spectrum_forward /
 * tntt_spectrum_inverse / tntt_polymul_spectrum / tntt_pointwise): spectra are [limbs][batch][n], canonical, in the
 * spectrum order of the plan's kernel shape; b_spectrum is [limbs][b_rows][n] with b_rows = batch or 1 (one spectrum
 * per limb shared by the whole batch). */
int tntt_rns_spectrum_forward(const tntt_rns_plan *plan, const void *in, void *out, size_t batch, void *cuda_stream);
int tntt_rns_spectrum_inverse(const tntt_rns_plan *plan, const void *in, void *out, size_t batch, void *cuda_stream);
int tntt_rns_polymul_spectrum(const tntt_rns_plan *plan, const void *a, const void *b_spectrum, void *c, size_t batch,
                              size_t b_rows, void *cuda_stream);
int tntt_rns_pointwise(const tntt_rns_plan *plan, const void *a, const void *b, void *c, size_t batch, void *cuda_stream);
/* test hook: regenerates limb `limb`'s tables with the host generators and compares them word for word with the
 * device-generated ones (TNTT_OK = identical) */
int tntt_rns_plan_check_tables(const tntt_rns_plan *plan, int limb);
int tntt_rns_kernel_attributes(const tntt_rns_plan *plan, int *regs, size_t *local_bytes, int *ctas_per_sm);
/* scripts/find_psi.py:9-44: the smallest psi in [2, max_search) with psi^n = -1 mod q (the script's own
 * max_search is 10000) -> returns 0.  If there is none that small the script gives up; this continues with
 * g^((q-1)/2n), g = 2, 3, ... and returns 1.  Host-side, like the script. */
int tntt_find_psi(uint32_t n, uint64_t q, uint64_t max_search, uint64_t *psi);

/* Integer-pipe microbenchmark (measurement only): kind 0 = IMAD.LO, 1 = IMAD.WIDE.U32, 2 = IADD3,
 * 3 = exact 64-bit Shoup modmul chain, 4 = 32-bit Shoup modmul chain, 5 = the 64-bit butterfly product
 * (approximate high product).  Returns thread-level ops per second. */
int tntt_microbench(int device, int kind, double *ops_per_second);

const char *tntt_last_error(void);
int tntt_version(void);
/* sizeof(tntt_plan_info) as the library was built: lets a binding check its mirror of the struct at load time */
size_t tntt_plan_info_size(void);

#ifdef __cplusplus
}
#endif
#endif /* TNTT_H */
