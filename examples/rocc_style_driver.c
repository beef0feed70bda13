/*
 * Plain-C client of libtntt.so that follows the command sequence of the reference's RoCC test
 * (chipyard/ntt-test.c:98-169): load polynomial A, load polynomial B, start, wait for done, read
 * the result, compare with the schoolbook product.  Host buffers in, host buffers out
 * (tntt_polymul_host); no Python, no torch.
 *
 *   gcc -O2 -Iinclude examples/rocc_style_driver.c -Ltiny-ntt_b200 -ltntt -Wl,-rpath,$PWD/tiny-ntt_b200 -o rocc_driver
 *   ./rocc_driver            # a = 1 + 2x + 3x^2, b = 5 + x  (the reference's vector), then a random batch
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tntt.h"

#define N 256u
#define Q 8380417ull
#define PSI 1239911ull

static void schoolbook(const uint32_t *a, const uint32_t *b, uint32_t *c) { /* ntt-test.c:60-80 */
    uint64_t acc[N];
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
    int rc = tntt_plan_create(&plan, 0, N, Q, PSI, 1);
    if (rc) { fprintf(stderr, "plan: %s\n", tntt_last_error()); return 2; }
    tntt_plan_info info;
    tntt_plan_info_get(plan, &info);
    if (info.word_bytes != 4) { fprintf(stderr, "expected 32-bit words\n"); return 2; }

    enum { ROWS = 64 };
    uint32_t *a = calloc(ROWS * N, 4), *b = calloc(ROWS * N, 4), *c = calloc(ROWS * N, 4), want[N];
    a[0] = 1; a[1] = 2; a[2] = 3;           /* ntt-test.c:101-107 */
    b[0] = 5; b[1] = 1;
    uint64_t x = 42;
    for (unsigned i = N; i < ROWS * N; ++i) {
        x = 6364136223846793005ULL * x + 1442695040888963407ULL;
        a[i] = (uint32_t)((x >> 17) % Q);
        x = 6364136223846793005ULL * x + 1442695040888963407ULL;
        b[i] = (uint32_t)((x >> 17) % Q);
    }
    rc = tntt_polymul_host(plan, a, b, c, ROWS);   /* load A, load B, start, poll done, read */
    if (rc) { fprintf(stderr, "polymul: %s\n", tntt_last_error()); return 2; }
    int bad = 0;
    for (unsigned r = 0; r < ROWS; ++r) {
        schoolbook(a + r * N, b + r * N, want);
        if (memcmp(want, c + r * N, sizeof want)) { ++bad; fprintf(stderr, "row %u differs\n", r); }
    }
    printf("c[0..3] = %u %u %u %u\n", c[0], c[1], c[2], c[3]);   /* 5 11 17 3 */
    printf(bad ? "FAIL\n" : "PASS\n");
    tntt_plan_destroy(plan);
    free(a); free(b); free(c);
    return bad ? 1 : 0;
}
