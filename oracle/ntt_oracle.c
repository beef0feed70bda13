/*
 * CPU oracle (plain C) for the tiny-ntt negacyclic-polymul path.
 * TEST INFRASTRUCTURE ONLY -- never linked into, loaded by, or called from the
 * product library (tiny-ntt_b200/csrc).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Parity status: PINNED by tests/test_oracle.py (golden vectors generated from
 * the reference's new_reference/cg_ntt.py, the reference KATs, and the C++
 * benchmark checksums of all four shipped parameter sets).
 *
 * It restates the reference golden model with explicit ring parameters:
 *   cg_ntt          new_reference/cg_ntt.py:29-65   (constant-geometry stages)
 *   cg_intt         new_reference/cg_ntt.py:68-75
 *   nwc_poly_mult   new_reference/cg_ntt.py:78-92
 *   make_poly/checksum  software_benchmark/benchmark_ntt.cpp:82-90,228-233 and
 *                       software_benchmark/benchmark_ntt_60bit.cpp:79-87,182-188
 * All words are uint64_t; q < 2^63.  Products use unsigned __int128 and `%`,
 * i.e. the reference arithmetic (benchmark_ntt_60bit.cpp:75-77), not Barrett.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

static inline uint64_t mulmod(uint64_t a, uint64_t b, uint64_t q) { return (uint64_t)(((u128)a * b) % q); }

static uint64_t powmod(uint64_t base, uint64_t e, uint64_t q) {
    uint64_t r = 1 % q;
    base %= q;
    while (e) {
        if (e & 1) r = mulmod(r, base, q);
        base = mulmod(base, base, q);
        e >>= 1;
    }
    return r;
}

static unsigned ilog2(uint32_t n) {
    unsigned l = 0;
    while ((1u << l) < n) ++l;
    return l;
}

static uint32_t bitrev(uint32_t v, unsigned bits) {
    uint32_t r = 0;
    for (unsigned i = 0; i < bits; ++i) r |= ((v >> i) & 1u) << (bits - 1 - i);
    return r;
}

uint64_t tntt_oracle_modinv(uint64_t v, uint64_t q) { return powmod(v, q - 2, q); } /* cg_ntt.py:9-10 */
uint64_t tntt_oracle_powmod(uint64_t b, uint64_t e, uint64_t q) { return powmod(b, e, q); }

/* cg_ntt.py:29-65.  `scratch` holds 2n words.  in == out allowed. */
static void cg_ntt_core(const uint64_t *in, uint64_t *out, uint32_t n, uint64_t omega, uint64_t q,
                        uint64_t *scratch) {
    const unsigned log_n = ilog2(n);
    const uint32_t half = n / 2;
    uint64_t *cur = scratch, *nxt = scratch + n;
    for (uint32_t i = 0; i < n; ++i) cur[bitrev(i, log_n)] = in[i] % q; /* :39, stores are % q (:57-59) */
    for (unsigned stage = 1; stage <= log_n; ++stage) {
        const uint32_t k = n >> stage;                 /* :50 */
        const uint64_t omega_s = powmod(omega, k, q);  /* :51 */
        uint64_t w = 1;
        for (uint32_t i = 0; i < half; ++i) {
            if (i && (i % k) == 0) w = mulmod(w, omega_s, q); /* omega_s^(i//k), :54 */
            const uint64_t left = cur[2 * i], t = mulmod(w, cur[2 * i + 1], q);
            uint64_t s = left + t;
            nxt[i] = s >= q ? s - q : s;               /* :58 */
            nxt[i + half] = left >= t ? left - t : left + q - t; /* :59 */
        }
        uint64_t *tmp = cur; cur = nxt; nxt = tmp;
    }
    memcpy(out, cur, (size_t)n * sizeof(uint64_t));
}

int tntt_oracle_cg_ntt(const uint64_t *in, uint64_t *out, uint32_t n, uint64_t omega, uint64_t q) {
    uint64_t *s = (uint64_t *)malloc((size_t)2 * n * sizeof(uint64_t));
    if (!s) return -1;
    cg_ntt_core(in, out, n, omega, q, s);
    free(s);
    return 0;
}

int tntt_oracle_cg_intt(const uint64_t *in, uint64_t *out, uint32_t n, uint64_t omega, uint64_t q) {
    uint64_t *s = (uint64_t *)malloc((size_t)2 * n * sizeof(uint64_t));
    if (!s) return -1;
    cg_ntt_core(in, out, n, powmod(omega, q - 2, q), q, s);      /* :72-73 */
    const uint64_t n_inv = powmod(n % q, q - 2, q);               /* :74 */
    for (uint32_t i = 0; i < n; ++i) out[i] = mulmod(out[i], n_inv, q);
    free(s);
    return 0;
}

/* cg_ntt.py:78-92; scratch holds 5n words */
static void nwc_core(const uint64_t *a, const uint64_t *b, uint64_t *c, uint32_t n, uint64_t psi, uint64_t q,
                     uint64_t *scratch) {
    uint64_t *fa = scratch, *fb = scratch + n, *tmp = scratch + 2 * n, *work = scratch + 3 * n;
    const uint64_t omega = mulmod(psi, psi, q);
    uint64_t p = 1;
    for (uint32_t i = 0; i < n; ++i) {                            /* :82-83 */
        fa[i] = mulmod(a[i] % q, p, q);
        fb[i] = mulmod(b[i] % q, p, q);
        p = mulmod(p, psi, q);
    }
    cg_ntt_core(fa, fa, n, omega, q, work);                       /* :86 */
    cg_ntt_core(fb, fb, n, omega, q, work);                       /* :87 */
    for (uint32_t i = 0; i < n; ++i) tmp[i] = mulmod(fa[i], fb[i], q); /* :88 */
    cg_ntt_core(tmp, tmp, n, powmod(omega, q - 2, q), q, work);   /* :90 */
    const uint64_t n_inv = powmod(n % q, q - 2, q), psi_inv = powmod(psi, q - 2, q);
    p = 1;
    for (uint32_t i = 0; i < n; ++i) {                            /* :74, :91-92 */
        c[i] = mulmod(mulmod(tmp[i], n_inv, q), p, q);
        p = mulmod(p, psi_inv, q);
    }
}

int tntt_oracle_nwc_poly_mult(const uint64_t *a, const uint64_t *b, uint64_t *c, uint32_t n, uint64_t psi,
                              uint64_t q) {
    uint64_t *s = (uint64_t *)malloc((size_t)5 * n * sizeof(uint64_t));
    if (!s) return -1;
    nwc_core(a, b, c, n, psi, q, s);
    free(s);
    return 0;
}

struct job { const uint64_t *a, *b; uint64_t *c; size_t lo, hi; uint32_t n; uint64_t psi, q; };

static void *worker(void *arg) {
    struct job *j = (struct job *)arg;
    uint64_t *s = (uint64_t *)malloc((size_t)5 * j->n * sizeof(uint64_t));
    for (size_t r = j->lo; r < j->hi; ++r)
        nwc_core(j->a + r * j->n, j->b + r * j->n, j->c + r * j->n, j->n, j->psi, j->q, s);
    free(s);
    return NULL;
}

/* rows of a [batch, n] array, split over `threads` host threads */
int tntt_oracle_nwc_poly_mult_batch(const uint64_t *a, const uint64_t *b, uint64_t *c, size_t batch, uint32_t n,
                                    uint64_t psi, uint64_t q, int threads) {
    if (threads < 1) threads = 1;
    if ((size_t)threads > batch) threads = batch ? (int)batch : 1;
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * threads);
    struct job *jobs = (struct job *)malloc(sizeof(struct job) * threads);
    for (int t = 0; t < threads; ++t) {
        jobs[t] = (struct job){a, b, c, batch * t / threads, batch * (t + 1) / threads, n, psi, q};
        pthread_create(&tid[t], NULL, worker, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(tid[t], NULL);
    free(tid);
    free(jobs);
    return 0;
}

/* O(n^2) definition: test_cg_ntt.py:11-21 */
void tntt_oracle_schoolbook(const uint64_t *a, const uint64_t *b, uint64_t *c, uint32_t n, uint64_t q) {
    memset(c, 0, (size_t)n * sizeof(uint64_t));
    for (uint32_t i = 0; i < n; ++i)
        for (uint32_t j = 0; j < n; ++j) {
            const uint64_t t = mulmod(a[i] % q, b[j] % q, q);
            const uint32_t d = i + j;
            if (d < n) { uint64_t s = c[d] + t; c[d] = s >= q ? s - q : s; }
            else { uint64_t v = c[d - n]; c[d - n] = v >= t ? v - t : v + q - t; }
        }
}

/* benchmark_ntt.cpp:82-90 (shift = 17) and benchmark_ntt_60bit.cpp:79-87 (shift = 0) */
void tntt_oracle_make_poly(uint64_t seed, uint32_t n, uint64_t q, int wide, uint64_t *out) {
    uint64_t x = seed;
    for (uint32_t i = 0; i < n; ++i) {
        x = 6364136223846793005ULL * x + 1442695040888963407ULL;
        out[i] = wide ? x % q : (x >> 17) % q;
    }
}

/* benchmark_ntt.cpp:228-233 (uint64 wrap, wide = 0); benchmark_ntt_60bit.cpp:182-188 (128-bit, wide = 1) */
uint64_t tntt_oracle_checksum(const uint64_t *v, uint32_t n, int wide) {
    uint64_t acc = 0;
    for (uint32_t i = 0; i < n; ++i)
        acc = wide ? (uint64_t)(((u128)acc * 1315423911ULL + v[i]) % 0xffffffffffffffc5ULL)
                   : (acc * 1315423911ULL + v[i]) % 0xffffffffffffffc5ULL;
    return acc;
}
