/*
 * The reference's C++ benchmark CLI (software_benchmark/benchmark_ntt.cpp:251-297, benchmark_ntt_60bit.cpp:205-251)
 * as a plain-C client of libtntt.so: same arguments (--check, --reps count), same inputs (make_poly(1), make_poly(2)),
 * same report -- target name, "N= Q= reps=", forward_ntt_total_ns / _avg_ns / _checksum, total_ns, avg_ns, checksum --
 * so a script that parses the reference binaries' output keeps working, and the two checksum lines are the ones the
 * reference prints for the same ring (e.g. checksum=2710933653778106521 for N=4096 / 60-bit).
 *
 * The reference fixes N, Q, PSI at configure time (-DBENCH_N ... in software_benchmark/CMakeLists.txt); here they are
 * run-time options with the reference's defaults, and --batch B times B rows per launch (row r = the same pair; the
 * reference binaries are single-polynomial, `avg_ns` stays "per polynomial product").  The 24-bit build draws
 * (x >> 17) % Q and folds its checksum with 64-bit wrap-around, the 60-bit build draws x % Q and folds in 128 bits
 * (benchmark_ntt.cpp:82-90,228-233 vs benchmark_ntt_60bit.cpp:79-87,182-188): chosen by the word size of the plan.
 * Timing is device-resident, like the reference's loop over arrays that stay in cache: `reps` launches between two
 * device synchronisations.  No Python, no torch.
 *
 *   gcc -O2 -Iinclude -I/usr/local/cuda/include examples/benchmark_ntt_gpu.c -Ltiny-ntt_b200 -ltntt \
 *       -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/tiny-ntt_b200 -o benchmark_ntt_gpu
 *   ./benchmark_ntt_gpu --check --reps 100                                   # N=256 Q=8380417 (the reference's default)
 *   ./benchmark_ntt_gpu --n 4096 --q 1152921504606830593 --psi 431606828070683274 --batch 32768 --reps 20
 */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "tntt.h"

#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)
#define TN(x) do { if ((x) != TNTT_OK) { fprintf(stderr, "%s: %s\n", #x, tntt_last_error()); return 2; } } while (0)

static uint64_t now_ns(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (uint64_t)t.tv_sec * 1000000000ull + (uint64_t)t.tv_nsec;
}

/* make_poly(seed): one LCG step per coefficient; `wide` = the 60-bit build's draw */
static void make_poly(uint64_t seed, uint32_t n, uint64_t q, int wide, uint64_t *out) {
    uint64_t x = seed;
    for (uint32_t i = 0; i < n; ++i) {
        x = 6364136223846793005ULL * x + 1442695040888963407ULL;
        out[i] = wide ? x % q : (x >> 17) % q;
    }
}

static uint64_t checksum(const uint64_t *v, uint32_t n, int wide) {
    const uint64_t mod = 0xffffffffffffffc5ULL, mul = 1315423911ULL;
    uint64_t acc = 0;
    for (uint32_t i = 0; i < n; ++i)
        acc = wide ? (uint64_t)(((unsigned __int128)acc * mul + v[i]) % mod) : (acc * mul + v[i]) % mod;
    return acc;
}

/* negacyclic_mul_reference: the O(N^2) product the reference's --check compares with */
static void schoolbook(const uint64_t *a, const uint64_t *b, uint64_t *c, uint32_t n, uint64_t q) {
    memset(c, 0, n * sizeof *c);
    for (uint32_t i = 0; i < n; ++i)
        for (uint32_t j = 0; j < n; ++j) {
            const uint64_t t = (uint64_t)((unsigned __int128)a[i] * b[j] % q);
            const uint32_t k = (i + j) & (n - 1);
            if (i + j < n) { c[k] += t; if (c[k] >= q) c[k] -= q; }
            else c[k] = c[k] >= t ? c[k] - t : c[k] + q - t;
        }
}

/* host words (uint64) <-> the plan's word size */
static void pack_rows(void *dst, const uint64_t *row, uint32_t n, size_t rows, int word_bytes) {
    for (size_t r = 0; r < rows; ++r)
        for (uint32_t i = 0; i < n; ++i)
            if (word_bytes == 4) ((uint32_t *)dst)[r * n + i] = (uint32_t)row[i];
            else ((uint64_t *)dst)[r * n + i] = row[i];
}
static void unpack_row(uint64_t *row, const void *src, uint32_t n, size_t r, int word_bytes) {
    for (uint32_t i = 0; i < n; ++i) row[i] = word_bytes == 4 ? ((const uint32_t *)src)[r * n + i] : ((const uint64_t *)src)[r * n + i];
}

int main(int argc, char **argv) {
    uint32_t n = 256;
    uint64_t q = 8380417ull, psi = 1239911ull;   /* the reference's configure-time defaults */
    long reps = 100;
    size_t batch = 1;
    int check = 0, device = 0;
    for (int i = 1; i < argc; ++i) {
        const char *a = argv[i];
        if (!strcmp(a, "--check")) check = 1;
        else if (!strcmp(a, "--reps") && i + 1 < argc) { reps = atol(argv[++i]); if (reps < 1) reps = 1; }
        else if (!strcmp(a, "--n") && i + 1 < argc) n = (uint32_t)strtoul(argv[++i], NULL, 0);
        else if (!strcmp(a, "--q") && i + 1 < argc) q = strtoull(argv[++i], NULL, 0);
        else if (!strcmp(a, "--psi") && i + 1 < argc) psi = strtoull(argv[++i], NULL, 0);
        else if (!strcmp(a, "--batch") && i + 1 < argc) { batch = (size_t)strtoull(argv[++i], NULL, 0); if (batch < 1) batch = 1; }
        else if (!strcmp(a, "--device") && i + 1 < argc) device = atoi(argv[++i]);
        else { fprintf(stderr, "usage: benchmark [--check] [--reps count] [--n N --q Q --psi PSI] [--batch rows] [--device d]\n"); return 2; }
    }

    tntt_plan *plan = NULL;
    TN(tntt_plan_create(&plan, device, n, q, psi, 1));
    tntt_plan_info info;
    TN(tntt_plan_info_get(plan, &info));
    const int wb = info.word_bytes, wide = wb == 8;
    const size_t row_bytes = (size_t)n * wb, bytes = batch * row_bytes;

    uint64_t *a = malloc(n * sizeof *a), *b = malloc(n * sizeof *b), *out = malloc(n * sizeof *out), *ref = malloc(n * sizeof *ref);
    void *ha = malloc(bytes), *hb = malloc(bytes), *hc = malloc(bytes);
    if (!a || !b || !out || !ref || !ha || !hb || !hc) { fprintf(stderr, "out of host memory\n"); return 2; }
    make_poly(1, n, q, wide, a);
    make_poly(2, n, q, wide, b);
    pack_rows(ha, a, n, batch, wb);
    pack_rows(hb, b, n, batch, wb);

    void *da, *db, *dc;
    CU(cudaSetDevice(device));
    CU(cudaMalloc(&da, bytes)); CU(cudaMalloc(&db, bytes)); CU(cudaMalloc(&dc, bytes));
    CU(cudaMemcpy(da, ha, bytes, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(db, hb, bytes, cudaMemcpyHostToDevice));

    if (check) {   /* every row of the batch against the schoolbook product */
        schoolbook(a, b, ref, n, q);
        TN(tntt_polymul(plan, da, db, dc, batch, NULL));
        CU(cudaMemcpy(hc, dc, bytes, cudaMemcpyDeviceToHost));
        for (size_t r = 0; r < batch; ++r) {
            unpack_row(out, hc, n, r, wb);
            if (memcmp(out, ref, n * sizeof *out)) { fprintf(stderr, "correctness check failed\n"); return 1; }
        }
    }

    /* warm-up (the first launch also loads the kernels), then the reference's two timed loops */
    TN(tntt_forward(plan, da, dc, batch, TNTT_TWIST, NULL));
    TN(tntt_polymul(plan, da, db, dc, batch, NULL));
    CU(cudaDeviceSynchronize());

    uint64_t t0 = now_ns();
    for (long r = 0; r < reps; ++r) TN(tntt_forward(plan, da, dc, batch, TNTT_TWIST, NULL));   /* forward_ntt_bench */
    CU(cudaDeviceSynchronize());
    const uint64_t fwd_ns = now_ns() - t0;
    CU(cudaMemcpy(hc, dc, row_bytes, cudaMemcpyDeviceToHost));
    unpack_row(out, hc, n, 0, wb);
    const uint64_t fwd_checksum = checksum(out, n, wide);

    t0 = now_ns();
    for (long r = 0; r < reps; ++r) TN(tntt_polymul(plan, da, db, dc, batch, NULL));           /* negacyclic_mul_ntt */
    CU(cudaDeviceSynchronize());
    const uint64_t ns = now_ns() - t0;
    CU(cudaMemcpy(hc, dc, bytes, cudaMemcpyDeviceToHost));
    unpack_row(out, hc, n, batch - 1, wb);

    const uint64_t products = (uint64_t)reps * batch;
    printf("benchmark_ntt_gpu\n");
    printf("N=%u Q=%llu reps=%ld\n", n, (unsigned long long)q, reps);
    printf("forward_ntt_total_ns=%llu\n", (unsigned long long)fwd_ns);
    printf("forward_ntt_avg_ns=%llu\n", (unsigned long long)(fwd_ns / products));
    printf("forward_ntt_checksum=%llu\n", (unsigned long long)fwd_checksum);
    printf("total_ns=%llu\n", (unsigned long long)ns);
    printf("avg_ns=%llu\n", (unsigned long long)(ns / products));
    printf("checksum=%llu\n", (unsigned long long)checksum(out, n, wide));
    if (batch > 1)   /* extra lines, after the reference's report */
        printf("batch=%zu\npolymul_per_s=%.0f\nforward_ntt_per_s=%.0f\n", batch, products * 1e9 / (double)ns, products * 1e9 / (double)fwd_ns);

    cudaFree(da); cudaFree(db); cudaFree(dc);
    tntt_plan_destroy(plan);
    free(a); free(b); free(out); free(ref); free(ha); free(hb); free(hc);
    return 0;
}
