/*
 * Plain-C client of the transform-domain entry points of libtntt.so: the RLWE pattern the reference's report
 * targets (reports/final-report.tex:571-610) -- one fixed polynomial s ("the key") multiplied with many
 * polynomials a_i.  The key is transformed once (tntt_spectrum_forward) and every product is
 * tntt_polymul_spectrum(a, s_hat), which skips one of the three transforms of nwc_poly_mult
 * (new_reference/cg_ntt.py:86-87).  Device memory is managed with the CUDA runtime directly; results are
 * compared with tntt_polymul (all three transforms) and with a schoolbook product.  No Python, no torch.
 *
 *   gcc -O2 -Iinclude -I/usr/local/cuda/include examples/cached_operand_driver.c -Ltiny-ntt_b200 -ltntt \
 *       -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/tiny-ntt_b200 -o cached_operand
 */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tntt.h"

#define N 1024u
#define Q 8380417ull
#define PSI 5548360ull /* rtl/twiddle_forward_1024.hex[1] */
#define ROWS 96u

#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)
#define TN(x) do { if ((x) != TNTT_OK) { fprintf(stderr, "%s: %s\n", #x, tntt_last_error()); return 2; } } while (0)

static void schoolbook(const uint32_t *a, const uint32_t *b, uint32_t *c) {
    static uint64_t acc[N];
    memset(acc, 0, sizeof acc);
    for (unsigned i = 0; i < N; ++i)
        for (unsigned j = 0; j < N; ++j) {
            const uint64_t t = (uint64_t)a[i] * b[j] % Q;
            if (i + j < N) acc[i + j] = (acc[i + j] + t) % Q;
            else acc[i + j - N] = (acc[i + j - N] + Q - t) % Q;
        }
    for (unsigned i = 0; i < N; ++i) c[i] = (uint32_t)acc[i];
}

int main(void) {
    tntt_plan *plan = NULL;
    TN(tntt_plan_create(&plan, 0, N, Q, PSI, 1));
    tntt_plan_info info;
    TN(tntt_plan_info_get(plan, &info));
    if (info.word_bytes != 4 || !info.spectrum) { fprintf(stderr, "expected a 32-bit plan with transform-domain kernels\n"); return 2; }

    const size_t row = N * sizeof(uint32_t);
    uint32_t *a = malloc(ROWS * row), *s = malloc(row), *c1 = malloc(ROWS * row), *c2 = malloc(ROWS * row), *sb = malloc(ROWS * row);
    uint64_t x = 7;
    for (unsigned i = 0; i < ROWS * N; ++i) { x = 6364136223846793005ULL * x + 1442695040888963407ULL; a[i] = (uint32_t)((x >> 17) % Q); }
    for (unsigned i = 0; i < N; ++i) { x = 6364136223846793005ULL * x + 1442695040888963407ULL; s[i] = (uint32_t)((x >> 17) % Q); }
    for (unsigned r = 0; r < ROWS; ++r) memcpy(sb + r * N, s, row);   /* the key repeated, for the plain product */

    void *da, *ds, *dshat, *dsb, *dc1, *dc2;
    CU(cudaMalloc(&da, ROWS * row)); CU(cudaMalloc(&ds, row)); CU(cudaMalloc(&dshat, row));
    CU(cudaMalloc(&dsb, ROWS * row)); CU(cudaMalloc(&dc1, ROWS * row)); CU(cudaMalloc(&dc2, ROWS * row));
    CU(cudaMemcpy(da, a, ROWS * row, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ds, s, row, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dsb, sb, ROWS * row, cudaMemcpyHostToDevice));

    TN(tntt_spectrum_forward(plan, ds, dshat, 1, NULL));                      /* once per key */
    TN(tntt_polymul_spectrum(plan, da, dshat, dc1, ROWS, 1, NULL));           /* every product: b_rows = 1, shared */
    TN(tntt_polymul(plan, da, dsb, dc2, ROWS, NULL));                         /* the same products, all three transforms */
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(c1, dc1, ROWS * row, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(c2, dc2, ROWS * row, cudaMemcpyDeviceToHost));

    int bad = memcmp(c1, c2, ROWS * row) != 0;
    uint32_t want[N];
    for (unsigned r = 0; r < 3; ++r) {
        schoolbook(a + r * N, s, want);
        if (memcmp(want, c1 + r * N, row)) { ++bad; fprintf(stderr, "row %u differs from the schoolbook product\n", r); }
    }
    /* and back: inverse of the key's spectrum is the key */
    TN(tntt_spectrum_inverse(plan, dshat, dc2, 1, NULL));
    CU(cudaMemcpy(c2, dc2, row, cudaMemcpyDeviceToHost));
    bad += memcmp(c2, s, row) != 0;
    printf("%u products with one cached operand, c[0][0..3] = %u %u %u %u\n", ROWS, c1[0], c1[1], c1[2], c1[3]);
    printf(bad ? "FAIL\n" : "PASS\n");
    cudaFree(da); cudaFree(ds); cudaFree(dshat); cudaFree(dsb); cudaFree(dc1); cudaFree(dc2);
    tntt_plan_destroy(plan);
    free(a); free(s); free(c1); free(c2); free(sb);
    return bad ? 1 : 0;
}
