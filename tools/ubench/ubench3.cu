// Round-2 microbenchmark: cost of one 64-bit butterfly (x, y) <- (x + w*y, x - w*y + G q) as a function of
//   PROD : how w*y mod q is formed (see prod<>)
//   TWSRC: 0 = twiddles are kernel parameters (constant bank / uniform registers), 1 = twiddles live in registers
//   XALU : extra independent ALU instructions per butterfly (is ALU work free next to the multiplier pipe?)
// Each thread runs passes of 4 stages over 16 registers, no memory traffic.  All products are real arithmetic for
// q = 2^60 - 2^14 + 1 (the Solinas forms use that shape), but nothing is checked here: timing only.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ubench3 ubench3.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ void unpack(uint64_t v, uint32_t &lo, uint32_t &hi) { asm("mov.b64 {%0,%1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t pack(uint32_t lo, uint32_t hi) { uint64_t v; asm("mov.b64 %0, {%1,%2};" : "=l"(v) : "r"(lo), "r"(hi)); return v; }

struct Tw { uint64_t w, wp, w2; uint32_t wp32, w2p; };   // w, floor(w 2^64/q), w 2^32 mod q, floor(w 2^32/q), floor(w2 2^32/q)
struct Mod { uint64_t q, nq, qg, zero; };

// approximate high product (drops y0*p0): h or h-1
__device__ __forceinline__ void hi3(uint32_t y0, uint32_t y1, uint32_t p0, uint32_t p1, uint32_t &h0, uint32_t &h1) {
    asm("{\n\t.reg .u32 s0, s1, c;\n\t"
        "mul.lo.u32 s0, %2, %5;\n\t"
        "mul.hi.u32 s1, %2, %5;\n\t"
        "mad.lo.cc.u32 s0, %3, %4, s0;\n\t"
        "madc.hi.cc.u32 s1, %3, %4, s1;\n\t"
        "addc.u32 c, 0, 0;\n\t"
        "mad.lo.cc.u32 %0, %3, %5, s1;\n\t"
        "madc.hi.u32 %1, %3, %5, c;\n\t}"
        : "=r"(h0), "=r"(h1) : "r"(y0), "r"(y1), "r"(p0), "r"(p1));
}

template <int PROD> __device__ __forceinline__ uint64_t prod(uint64_t y, const Tw &t, const Mod &m) {
    uint32_t y0, y1, w0, w1, p0, p1, n0, n1, h0, h1, lo, hi;
    unpack(y, y0, y1); unpack(t.w, w0, w1); unpack(t.wp, p0, p1); unpack(m.nq, n0, n1);
    uint64_t acc;
    if constexpr (PROD == 0) {            // the round-1 product: 5 half-rate + 4 full-rate multiplies
        hi3(y0, y1, p0, p1, h0, h1);
        asm("mul.wide.u32 %0, %1, %2;" : "=l"(acc) : "r"(y0), "r"(w0));
        asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(h0), "r"(n0));
        unpack(acc, lo, hi);
        asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(h0), "r"(n1));
        asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(h1), "r"(n0));
        asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(y0), "r"(w1));
        asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(y1), "r"(w0));
        return pack(lo, hi);
    } else if constexpr (PROD == 1) {     // split Shoup: y1*w2 + y0*w - h q with a 33-bit quotient: 5 half + 3 full + 1 predicated add
        uint32_t v0, v1, c;
        unpack(t.w2, v0, v1);
        asm("{\n\t.reg .u32 s0, s1;\n\t"
            "mul.lo.u32 s0, %2, %4;\n\t"
            "mul.hi.u32 s1, %2, %4;\n\t"
            "mad.lo.cc.u32 s0, %3, %5, s0;\n\t"
            "madc.hi.cc.u32 %0, %3, %5, s1;\n\t"
            "addc.u32 %1, 0, 0;\n\t}"
            : "=r"(h0), "=r"(c) : "r"(y1), "r"(y0), "r"(t.w2p), "r"(t.wp32));
        asm("mul.wide.u32 %0, %1, %2;" : "=l"(acc) : "r"(y0), "r"(w0));
        asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(y1), "r"(v0));
        asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(h0), "r"(n0));
        unpack(acc, lo, hi);
        asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(h0), "r"(n1));
        asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(y0), "r"(w1));
        asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(y1), "r"(v1));
        hi += c ? n0 : 0u;
        return pack(lo, hi);
    } else if constexpr (PROD == 2) {     // Shoup, h*nq entirely on the ALU for q = 2^60 - 2^14 + 1: h*nq = (h << 14) - h - (h << 60)
        hi3(y0, y1, p0, p1, h0, h1);
        asm("mul.wide.u32 %0, %1, %2;" : "=l"(acc) : "r"(y0), "r"(w0));
        unpack(acc, lo, hi);
        asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(y0), "r"(w1));
        asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(y1), "r"(w0));
        uint32_t a14, f, t28;
        asm("shf.l.wrap.b32 %0, %1, %2, 14;" : "=r"(a14) : "r"((uint32_t)m.zero), "r"(h0));   // h0 << 14
        asm("shf.l.wrap.b32 %0, %1, %2, 14;" : "=r"(f) : "r"(h0), "r"(h1));                     // (h << 14) >> 32
        asm("shf.l.wrap.b32 %0, %1, %2, 28;" : "=r"(t28) : "r"((uint32_t)m.zero), "r"(h0));     // h0 << 28
        asm("{\n\t.reg .u32 u;\n\t"
            "add.cc.u32 %0, %0, %2;\n\t"       // lo += h0 << 14
            "addc.u32 %1, %1, %3;\n\t"         // hi += f + carry
            "sub.cc.u32 %0, %0, %4;\n\t"       // lo -= h0
            "subc.u32 %1, %1, %5;\n\t"         // hi -= h1 + borrow
            "sub.u32 %1, %1, %6;\n\t}"         // hi -= h0 << 28
            : "+r"(lo), "+r"(hi) : "r"(a14), "r"(f), "r"(h0), "r"(h1), "r"(t28));
        return pack(lo, hi);
    } else if constexpr (PROD == 3) {     // Shoup with only the cross terms h0*n1 + h1*n0 on the ALU
        hi3(y0, y1, p0, p1, h0, h1);
        asm("mul.wide.u32 %0, %1, %2;" : "=l"(acc) : "r"(y0), "r"(w0));
        asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(h0), "r"(n0));
        unpack(acc, lo, hi);
        asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(y0), "r"(w1));
        asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(y1), "r"(w0));
        uint32_t t1, t2;
        const uint32_t z = (uint32_t)m.zero;
        asm("shf.l.wrap.b32 %0, %1, %2, 14;" : "=r"(t1) : "r"(z), "r"(h1));
        asm("shf.l.wrap.b32 %0, %1, %2, 28;" : "=r"(t2) : "r"(z), "r"(h0));
        hi = hi + t1 - h1;
        hi = hi - t2 + z;
        return pack(lo, hi);
    } else {                              // PROD == 4: full 124-bit product, then 2^60 = 2^14 - 1 folds on the ALU (no quotient at all)
        uint64_t t0, t1, t2, t3;
        asm("mul.wide.u32 %0, %1, %2;" : "=l"(t0) : "r"(y0), "r"(w0));
        uint32_t t0l, t0h; unpack(t0, t0l, t0h);
        asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t1) : "r"(y0), "r"(w1), "l"((uint64_t)t0h));
        uint32_t t1l, t1h; unpack(t1, t1l, t1h);
        asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t2) : "r"(y1), "r"(w0), "l"((uint64_t)t1l));
        uint32_t t2l, t2h; unpack(t2, t2l, t2h);
        asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t3) : "r"(y1), "r"(w1), "l"((uint64_t)t1h));
        t3 += t2h;
        uint32_t p2, p3; unpack(t3, p2, p3);           // product = p3:p2:t2l:t0l  (< 2^124)
        // A = P mod 2^60, B = P >> 60 (64 bits); P = A + B (2^14 - 1)
        const uint32_t a1 = t2l & 0x0FFFFFFFu;
        uint32_t b0, b1;
        asm("shf.r.wrap.b32 %0, %1, %2, 28;" : "=r"(b0) : "r"(t2l), "r"(p2));
        asm("shf.r.wrap.b32 %0, %1, %2, 28;" : "=r"(b1) : "r"(p2), "r"(p3));
        uint32_t c0, c1, c2;
        asm("shf.l.wrap.b32 %0, %1, %2, 14;" : "=r"(c0) : "r"((uint32_t)m.zero), "r"(b0));
        asm("shf.l.wrap.b32 %0, %1, %2, 14;" : "=r"(c1) : "r"(b0), "r"(b1));
        asm("shf.r.wrap.b32 %0, %1, %2, 18;" : "=r"(c2) : "r"(b1), "r"((uint32_t)m.zero));
        uint32_t s0, s1, s2;                            // S = A + (B << 14) - B  (< 2^79)
        asm("{\n\t"
            "add.cc.u32 %0, %3, %5;\n\t"
            "addc.cc.u32 %1, %4, %6;\n\t"
            "addc.u32 %2, %7, 0;\n\t"
            "sub.cc.u32 %0, %0, %8;\n\t"
            "subc.cc.u32 %1, %1, %9;\n\t"
            "subc.u32 %2, %2, 0;\n\t}"
            : "=&r"(s0), "=&r"(s1), "=&r"(s2) : "r"(t0l), "r"(a1), "r"(c0), "r"(c1), "r"(c2), "r"(b0), "r"(b1));
        // second fold: S = A' + B' 2^60, B' < 2^19
        const uint32_t a1b = s1 & 0x0FFFFFFFu;
        uint32_t bb;
        asm("shf.r.wrap.b32 %0, %1, %2, 28;" : "=r"(bb) : "r"(s1), "r"(s2));
        uint32_t d0 = bb << 14, d1 = bb >> 18;
        asm("{\n\t"
            "add.cc.u32 %0, %2, %4;\n\t"
            "addc.u32 %1, %3, %5;\n\t"
            "sub.cc.u32 %0, %0, %6;\n\t"
            "subc.u32 %1, %1, 0;\n\t}"
            : "=&r"(lo), "=&r"(hi) : "r"(s0), "r"(a1b), "r"(d0), "r"(d1), "r"(bb));
        return pack(lo, hi);
    }
}

template <int PROD, int XALU>
__device__ __forceinline__ void bfly(uint64_t &x, uint64_t &y, const Tw &t, const Mod &m, uint32_t &dummy) {
    const uint64_t v = prod<PROD>(y, t, m);
    y = x - v + m.qg;
    x = x + v + m.zero;
#pragma unroll
    for (int i = 0; i < XALU; ++i) asm("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(dummy) : "r"((uint32_t)x), "r"((uint32_t)y));
}

struct Params { Tw tw[4]; Mod mod; };

template <int PROD, int TWSRC, int XALU, int MINB>
__global__ void __launch_bounds__(256, MINB) chain(uint64_t *sink, int iters, const __grid_constant__ Params P) {
    uint64_t x[16];
    uint32_t dummy = threadIdx.x;
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = (uint64_t)threadIdx.x * 0x9E3779B97F4A7C15ull + k * 0x1234567ull;
    Tw tw[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        tw[s] = P.tw[s];
        if (TWSRC == 1) {   // make them thread-dependent so that they must live in registers
            tw[s].w += threadIdx.x; tw[s].wp ^= threadIdx.x; tw[s].w2 += threadIdx.x; tw[s].wp32 ^= threadIdx.x; tw[s].w2p += threadIdx.x;
        }
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const int kb = 3 - s;
#pragma unroll
            for (int g = 0; g < (16 >> (kb + 1)); ++g)
#pragma unroll
                for (int j = 0; j < (1 << kb); ++j) {
                    const int k0 = (g << (kb + 1)) | j, k1 = k0 | (1 << kb);
                    bfly<PROD, XALU>(x[k0], x[k1], tw[s], P.mod, dummy);
                }
        }
    }
    uint64_t r = dummy;
#pragma unroll
    for (int k = 0; k < 16; ++k) r ^= x[k];
    if (r == 0x12345678ull) sink[0] = r;
}

template <int PROD, int TWSRC, int XALU, int MINB> void run(const char *name, uint64_t *sink, const Params &P) {
    auto kern = chain<PROD, TWSRC, XALU, MINB>;
    const int smem = MINB >= 4 ? 48 * 1024 : (MINB == 3 ? 72 * 1024 : (MINB == 2 ? 100 * 1024 : 200 * 1024));
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int bps = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, 256, smem));
    cudaFuncAttributes attr; CK(cudaFuncGetAttributes(&attr, kern));
    const int iters = 256, blocks = 148 * bps * 2;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        kern<<<blocks, 256, smem>>>(sink, iters, P);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const double bf = (double)blocks * 256 * iters * 32 / (best * 1e-3);
    printf("%-46s regs=%3d local=%3zu warps/SMSP=%d  %6.2f SMSP-cycles per warp-butterfly\n", name, attr.numRegs,
           (size_t)attr.localSizeBytes, bps * 2, 148.0 * 4 * 1.965e9 / (bf / 32));
}

int main() {
    const uint64_t Q = 1152921504606830593ull;
    Params P;
    const uint64_t ws[4] = {431606828070683274ull, 164227591873870967ull, 1152640029630119941ull, 99887766554433221ull};
    for (int s = 0; s < 4; ++s) {
        const unsigned __int128 w = ws[s];
        P.tw[s].w = ws[s];
        P.tw[s].wp = (uint64_t)((w << 64) / Q);
        P.tw[s].w2 = (uint64_t)((w << 32) % Q);
        P.tw[s].wp32 = (uint32_t)((w << 32) / Q);
        P.tw[s].w2p = (uint32_t)((((unsigned __int128)P.tw[s].w2) << 32) / Q);
    }
    P.mod.q = Q; P.mod.nq = 0 - Q; P.mod.qg = 3 * Q; P.mod.zero = 0;
    uint64_t *sink; CK(cudaMalloc(&sink, 64));
#define RUN(P_, T_, X_, B_) run<P_, T_, X_, B_>("prod" #P_ " twsrc" #T_ " xalu" #X_ " b" #B_, sink, P)
    printf("== product 0 (round 1): twiddle source, extra ALU work, occupancy\n");
    RUN(0, 0, 0, 3); RUN(0, 1, 0, 3); RUN(0, 0, 0, 2); RUN(0, 1, 0, 2); RUN(0, 1, 0, 4);
    RUN(0, 0, 1, 3); RUN(0, 0, 2, 3); RUN(0, 0, 4, 3); RUN(0, 1, 1, 3); RUN(0, 1, 2, 3); RUN(0, 1, 4, 3);
    printf("== product 1 (split Shoup, 33-bit quotient)\n");
    RUN(1, 0, 0, 3); RUN(1, 1, 0, 3); RUN(1, 1, 0, 2);
    printf("== product 2 (Shoup, h*nq on the ALU)\n");
    RUN(2, 0, 0, 3); RUN(2, 1, 0, 3); RUN(2, 1, 0, 2);
    printf("== product 3 (Shoup, cross terms of h*nq on the ALU)\n");
    RUN(3, 0, 0, 3); RUN(3, 1, 0, 3); RUN(3, 1, 0, 2);
    printf("== product 4 (full product + Solinas folds on the ALU)\n");
    RUN(4, 0, 0, 3); RUN(4, 1, 0, 3); RUN(4, 1, 0, 2);
    return 0;
}
