/*
 * Plain-C client of the multi-modulus entry points of libtntt.so: the RNS / FHE use the reference's report names
 * as the step after its single 60-bit modulus (reports/final-report.tex:1811-1817).
 *
 *   - three NTT-friendly primes q_l = 1 (mod 2N) and their roots psi_l: tntt_find_psi, the library form of
 *     scripts/find_psi.py:9-44
 *   - one plan for all limbs, twiddle / Shoup tables generated on the device: tntt_rns_plan_create
 *   - [L][B][N] residues multiplied limb by limb in ONE launch: tntt_rns_polymul
 *   - one operand kept in the transform domain: tntt_rns_spectrum_forward + tntt_rns_polymul_spectrum
 *   - the same product for the first limb through host buffers on every GPU of the box: tntt_polymul_host_multi
 *
 * Row 0 of every limb is compared with a schoolbook negacyclic product (new_reference/cg_ntt.py:78-92 computes the
 * same ring product); everything else is compared between the entry points.  No Python, no torch.
 *
 *   gcc -O2 -Iinclude -I/usr/local/cuda/include examples/rns_driver.c -Ltiny-ntt_b200 -ltntt \
 *       -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/tiny-ntt_b200 -o rns_driver
 */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tntt.h"

#define N 256u
#define LIMBS 3
#define ROWS 40u

#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)
#define TN(x) do { if ((x) < TNTT_OK) { fprintf(stderr, "%s: %s\n", #x, tntt_last_error()); return 2; } } while (0)

typedef unsigned __int128 u128;

static void schoolbook(const uint64_t *a, const uint64_t *b, uint64_t *c, uint64_t q) {
    memset(c, 0, N * sizeof(uint64_t));
    for (unsigned i = 0; i < N; ++i)
        for (unsigned j = 0; j < N; ++j) {
            const uint64_t t = (uint64_t)((u128)a[i] * b[j] % q);
            if (i + j < N) c[i + j] = (c[i + j] + t) % q;
            else c[i + j - N] = (c[i + j - N] + q - t) % q;
        }
}

int main(void) {
    /* primes just below 2^58 with q = 1 (mod 512); tntt_find_psi rejects every candidate that is not a prime of that form */
    uint64_t q[LIMBS], psi[LIMBS];
    int found = 0;
    for (uint64_t c = (1ull << 58) - 511; found < LIMBS && c > (1ull << 57); c -= 2 * N)
        if (tntt_find_psi(N, c, 10000, &psi[found]) >= 0) q[found++] = c;
    if (found < LIMBS) { fprintf(stderr, "no primes found\n"); return 2; }

    tntt_rns_plan *plan = NULL;
    TN(tntt_rns_plan_create(&plan, 0, N, q, psi, LIMBS));
    printf("rns plan: %d limbs, %d-byte words, kernel %s, %zu bytes of device-generated tables\n", tntt_rns_plan_limbs(plan),
           tntt_rns_plan_word_bytes(plan), tntt_rns_plan_kernel(plan), tntt_rns_plan_table_bytes(plan));
    for (int l = 0; l < LIMBS; ++l) TN(tntt_rns_plan_check_tables(plan, l));     /* == the host generators, word for word */

    const size_t limb_words = (size_t)ROWS * N, words = LIMBS * limb_words, bytes = words * sizeof(uint64_t);
    uint64_t *a = malloc(bytes), *b = malloc(bytes), *c = malloc(bytes), *c2 = malloc(bytes), *want = malloc(N * sizeof(uint64_t));
    uint64_t x = 11;
    for (int l = 0; l < LIMBS; ++l)
        for (size_t i = 0; i < limb_words; ++i) {
            x = 6364136223846793005ULL * x + 1442695040888963407ULL; a[l * limb_words + i] = (x >> 3) % q[l];
            x = 6364136223846793005ULL * x + 1442695040888963407ULL; b[l * limb_words + i] = (x >> 3) % q[l];
        }
    void *da, *db, *dbhat, *dc, *dc2;
    CU(cudaMalloc(&da, bytes)); CU(cudaMalloc(&db, bytes)); CU(cudaMalloc(&dbhat, bytes)); CU(cudaMalloc(&dc, bytes)); CU(cudaMalloc(&dc2, bytes));
    CU(cudaMemcpy(da, a, bytes, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(db, b, bytes, cudaMemcpyHostToDevice));

    TN(tntt_rns_polymul(plan, da, db, dc, ROWS, NULL));                              /* all limbs, one launch */
    TN(tntt_rns_spectrum_forward(plan, db, dbhat, ROWS, NULL));                      /* b kept in the transform domain */
    TN(tntt_rns_polymul_spectrum(plan, da, dbhat, dc2, ROWS, ROWS, NULL));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(c, dc, bytes, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(c2, dc2, bytes, cudaMemcpyDeviceToHost));
    int bad = memcmp(c, c2, bytes) != 0;
    for (int l = 0; l < LIMBS; ++l) {
        schoolbook(a + l * limb_words, b + l * limb_words, want, q[l]);
        bad |= memcmp(want, c + l * limb_words, N * sizeof(uint64_t)) != 0;
    }

    /* limb 0 once more through host buffers, sharded over every visible GPU by one call */
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (ndev > 8) ndev = 8;
    tntt_plan *plans[8];
    for (int d = 0; d < ndev; ++d) TN(tntt_plan_create(&plans[d], d, N, q[0], psi[0], 1));
    void *ha, *hb, *hc;
    CU(cudaMallocHost(&ha, limb_words * 8)); CU(cudaMallocHost(&hb, limb_words * 8)); CU(cudaMallocHost(&hc, limb_words * 8));
    memcpy(ha, a, limb_words * 8);
    memcpy(hb, b, limb_words * 8);
    TN(tntt_polymul_host_multi(plans, ndev, ha, hb, hc, ROWS));
    bad |= memcmp(hc, c, limb_words * 8) != 0;
    printf("limb 0 over %d GPU(s) through host buffers: %s\n", ndev, memcmp(hc, c, limb_words * 8) ? "MISMATCH" : "same bits");

    for (int d = 0; d < ndev; ++d) tntt_plan_destroy(plans[d]);
    tntt_rns_plan_destroy(plan);
    printf("c[limb 0][row 0][0..3] = %llu %llu %llu %llu\n", (unsigned long long)c[0], (unsigned long long)c[1], (unsigned long long)c[2],
           (unsigned long long)c[3]);
    printf(bad ? "FAIL\n" : "PASS\n");
    return bad;
}
