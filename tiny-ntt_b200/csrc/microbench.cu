// Integer-pipe microbenchmarks: the IMAD roofline denominator is measured, not assumed
// (MEASURED_PEAKS.json has no integer-pipe figure).
#include "common.h"

namespace tntt {

constexpr int kIters = 4096, kIlp = 8;

template <int KIND> __global__ void __launch_bounds__(256) intpipe_kernel(uint32_t *sink, uint32_t seed) {
    uint32_t a[kIlp], b[kIlp];
    uint64_t w[kIlp];
#pragma unroll
    for (int i = 0; i < kIlp; ++i) { a[i] = seed + threadIdx.x + i; b[i] = seed * 3 + i; w[i] = a[i]; }
    const uint32_t m = seed | 1u;
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < kIlp; ++i) {
            if (KIND == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(b[i]));
            if (KIND == 1) {  // operands change every iteration, otherwise ptxas hoists the product out of the loop
                uint32_t lo, hi;
                asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(w[i]));
                asm("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(lo), "r"(hi));
            }
            if (KIND == 2) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < kIlp; ++i) r ^= a[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    if (r == 0x12345678u) sink[0] = r;
}

template <typename W, bool LAZY> __global__ void __launch_bounds__(256) modmul_kernel(W *sink, W seed, Mod<W> mod, Tw<W> t) {
    W x[kIlp];
#pragma unroll
    for (int i = 0; i < kIlp; ++i) x[i] = seed + threadIdx.x * 977u + i;
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < kIlp; ++i) x[i] = LAZY ? shoup_lazy(x[i], t.w, t.wp, mod) : shoup_mul(x[i], t.w, t.wp, mod.nq);
    }
    W r = 0;
#pragma unroll
    for (int i = 0; i < kIlp; ++i) r ^= x[i];
    if (r == (W)0x12345678u) sink[0] = r;
}

cudaError_t run_microbench(int kind, double *ops_per_second) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    void *sink = nullptr;
    cudaError_t e = cudaMalloc(&sink, 64);
    if (e != cudaSuccess) return e;
    const int blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    Mod<uint64_t> m64{}; m64.q = 1152921504606830593ull; m64.nq = 0 - m64.q; m64.qg = 3 * m64.q;
    Mod<uint32_t> m32{}; m32.q = 8380417u; m32.nq = 0u - m32.q;
    Tw<uint64_t> t64{431606828070683274ull, 6905709249130932383ull};
    Tw<uint32_t> t32{1239911u, 635448320u};
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        switch (kind) {
            case 0: intpipe_kernel<0><<<blocks, threads>>>((uint32_t *)sink, 12345u + rep); break;
            case 1: intpipe_kernel<1><<<blocks, threads>>>((uint32_t *)sink, 12345u + rep); break;
            case 2: intpipe_kernel<2><<<blocks, threads>>>((uint32_t *)sink, 12345u + rep); break;
            case 3: modmul_kernel<uint64_t, false><<<blocks, threads>>>((uint64_t *)sink, 99ull + rep, m64, t64); break;
            case 5: modmul_kernel<uint64_t, true><<<blocks, threads>>>((uint64_t *)sink, 99ull + rep, m64, t64); break;
            case 4: modmul_kernel<uint32_t, false><<<blocks, threads>>>((uint32_t *)sink, 99u + rep, m32, t32); break;
            default: cudaFree(sink); return cudaErrorInvalidValue;
        }
        cudaEventRecord(e1);
        e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    if (e != cudaSuccess) return e;
    *ops_per_second = (double)blocks * threads * (double)kIters * kIlp / (best * 1e-3);
    return cudaGetLastError();
}

}  // namespace tntt
