// Throughput microbenchmarks of the integer instruction sequences the NTT butterflies are made of.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu ; run on the B200.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 2048, ILP = 8;

__device__ __forceinline__ void unpack(uint64_t v, uint32_t& lo, uint32_t& hi){ asm("mov.b64 {%0,%1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t pack(uint32_t lo, uint32_t hi){ uint64_t v; asm("mov.b64 %0, {%1,%2};" : "=l"(v) : "r"(lo), "r"(hi)); return v; }

// mulhi without carry flags: zero-extended 32-bit addends only
__device__ __forceinline__ uint64_t mulhi_nc(uint64_t a, uint64_t b) {
    uint32_t a0, a1, b0, b1; unpack(a, a0, a1); unpack(b, b0, b1);
    uint64_t t0, t1, t2, h;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(t0) : "r"(a0), "r"(b0));
    uint32_t t0l, t0h; unpack(t0, t0l, t0h);
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t1) : "r"(a0), "r"(b1), "l"((uint64_t)t0h));
    uint32_t t1l, t1h; unpack(t1, t1l, t1h);
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t2) : "r"(a1), "r"(b0), "l"((uint64_t)t1l));
    uint32_t t2l, t2h; unpack(t2, t2l, t2h);
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(h) : "r"(a1), "r"(b1), "l"((uint64_t)t1h));
    return h + t2h;
}
__device__ __forceinline__ uint64_t tpart(uint64_t x, uint64_t w, uint64_t h, uint64_t nq) {
    uint32_t x0,x1,w0,w1,h0,h1,n0,n1,lo,hi; unpack(x,x0,x1); unpack(w,w0,w1); unpack(h,h0,h1); unpack(nq,n0,n1);
    uint64_t acc;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(acc) : "r"(x0), "r"(w0));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(h0), "r"(n0));
    unpack(acc, lo, hi);
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(x0), "r"(w1));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(x1), "r"(w0));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(h0), "r"(n1));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(h1), "r"(n0));
    return pack(lo, hi);
}

template <int KIND> __global__ void __launch_bounds__(256) k(uint64_t* sink, uint64_t seed, uint64_t w, uint64_t wp, uint64_t nq) {
    uint64_t x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = seed + threadIdx.x * 977u + i * 1315423911ull;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (KIND == 0) x[i] = __umul64hi(x[i], wp);                       // compiler mulhi (carry-flag forms)
            if (KIND == 1) x[i] = mulhi_nc(x[i], wp);                          // carry-free mulhi
            if (KIND == 2) x[i] = tpart(x[i], w, x[i] >> 3, nq);               // 6-IMAD low part
            if (KIND == 3) { uint64_t h = __umul64hi(x[i], wp); x[i] = tpart(x[i], w, h, nq); }   // full Shoup
            if (KIND == 4) { uint64_t h = mulhi_nc(x[i], wp); x[i] = tpart(x[i], w, h, nq); }     // Shoup, carry-free mulhi
            if (KIND == 5) x[i] = x[i] * w;                                    // mul.lo.u64
            if (KIND == 6) x[i] = x[i] + w + (x[i] >> 63);                     // 64-bit adds
            if (KIND == 7) { uint32_t lo, hi; unpack(x[i], lo, hi); uint64_t t; asm("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(lo), "r"(hi)); x[i] = t; }
            if (KIND == 9) { uint32_t lo, hi; unpack(x[i], lo, hi); asm("mul.hi.u32 %0, %0, %1;" : "+r"(lo) : "r"(hi | 0x80000001u)); x[i] = pack(lo, hi); }
            if (KIND == 10) { uint32_t lo, hi; unpack(x[i], lo, hi); asm("mad.hi.u32 %0, %0, %1, %1;" : "+r"(lo) : "r"(hi | 0x80000001u)); x[i] = pack(lo, hi); }
            if (KIND == 8) { uint32_t lo, hi; unpack(x[i], lo, hi); asm("mad.lo.u32 %0, %0, %1, %1;" : "+r"(lo) : "r"(hi)); x[i] = pack(lo, hi); }
        }
    }
    uint64_t r = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) r ^= x[i];
    if (r == 0x12345678ull) sink[0] = r;
}

template <int KIND> double run(const char* name, int ctas_per_sm, double ops_per_item) {
    uint64_t* sink; cudaMalloc(&sink, 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * ctas_per_sm;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<KIND><<<blocks, 256>>>(sink, 99 + rep, 431606828070683274ull, 6905709249130932383ull, 0ull - 1152921504606830593ull);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const double items = (double)blocks * 256 * ITERS * ILP / (best * 1e-3);
    const double cyc = 148.0 * 4 * 1.965e9 / (items / 32);
    printf("%-34s ctas/SM=%d  %.3e items/s  %.2f SMSP-cycles per warp-item\n", name, ctas_per_sm, items, cyc);
    cudaFree(sink);
    return items;
}

int main() {
    for (int c : {8}) {
        run<0>("umul64hi (compiler)", c, 1);
        run<1>("mulhi carry-free", c, 1);
        run<2>("T part (6 IMAD)", c, 1);
        run<3>("Shoup = umul64hi + T", c, 1);
        run<4>("Shoup = carry-free mulhi + T", c, 1);
        run<5>("mul.lo.u64", c, 1);
        run<6>("64-bit add x2", c, 1);
        run<7>("mul.wide.u32", c, 1);
        run<8>("mad.lo.u32", c, 1);
        run<9>("mul.hi.u32", c, 1);
        run<10>("mad.hi.u32", c, 1);
    }
    return 0;
}
