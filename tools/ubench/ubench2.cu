// Round-2 microbenchmarks: what bounds the fused 64-bit kernel besides the multiplier pipe?
//  (1) the kernel's own register passes (fwd_pass of kernels.cuh) with no memory traffic, at 1..4 CTAs per SM:
//      how much of the multiplier pipe can 2 / 4 / 6 / 8 warps per SMSP keep busy?
//  (2) IMAD.WIDE / IMAD.LO streams with k ALU instructions (LOP3 / SHF / IADD3) beside each multiply:
//      do ALU instructions issue for free next to a saturated multiplier pipe?
//  (3) IMAD.WIDE with register, constant-bank and uniform-register multiplicands (operand delivery).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ubench2 ubench2.cu ; run on the B200.
#include <cstdint>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>

#include "../../tiny-ntt_b200/csrc/common.h"
#include "../../tiny-ntt_b200/csrc/tables.h"

using namespace tntt;
using W = uint64_t;
using C = Cfg<W, 12, 4, 1>;
constexpr uint64_t Q = 1152921504606830593ull, PSI = 431606828070683274ull;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

// ------------------------------------------------------------------ (1) register passes only
template <int MINB, int MODE>
__global__ void __launch_bounds__(256, MINB)
pass_chain(W *sink, int iters, const __grid_constant__ PolymulTables<W> tb, const __grid_constant__ Mod<W> mod) {
    const int tid = threadIdx.x;
    W x[1][C::R];
#pragma unroll
    for (int k = 0; k < C::R; ++k) x[0][k] = (W)tid * 0x9E3779B97F4A7C15ull + k * 0x1234567ull;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 3) fwd_pass<C, 0, 1, true>(x, tid, tb, mod);          // uniform twiddles (constant bank)
        if (MODE == 1 || MODE == 3) fwd_pass<C, 1, 1, true>(x, tid, tb, mod);          // broadcast twiddle loads
        if (MODE == 2 || MODE == 3) fwd_pass<C, 2, 1, true>(x, tid, tb, mod);          // per-thread twiddle loads
    }
    W r = 0;
#pragma unroll
    for (int k = 0; k < C::R; ++k) r ^= x[0][k];
    if (r == 0x12345678ull) sink[0] = r;
}

template <int MINB, int MODE> void run_pass(const char *name, int ctas_per_sm, const PolymulTables<W> &tb, const Mod<W> &mod, W *sink) {
    auto kern = pass_chain<MINB, MODE>;
    // dynamic shared memory pins the number of resident CTAs per SM
    const int smem = ctas_per_sm >= 4 ? 48 * 1024 : (ctas_per_sm == 3 ? 72 * 1024 : (ctas_per_sm == 2 ? 100 * 1024 : 200 * 1024));
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int bps = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, 256, smem));
    cudaFuncAttributes attr; CK(cudaFuncGetAttributes(&attr, kern));
    const int iters = 64, blocks = 148 * bps * 4;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        kern<<<blocks, 256, smem>>>(sink, iters, tb, mod);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const int passes = MODE == 3 ? 3 : 1;
    const double bf = (double)blocks * 256 * iters * passes * 32 / (best * 1e-3);   // butterflies per second (32 per thread per pass)
    const double cyc = 148.0 * 4 * 1.965e9 / (bf / 32);
    printf("%-30s regs=%3d local=%3zu ctas/SM=%d (%d warps/SMSP)  %.3e butterflies/s  %.2f SMSP-cycles per warp-butterfly (28 = multiplier pipe full)\n",
           name, attr.numRegs, (size_t)attr.localSizeBytes, bps, bps * 2, bf, cyc);
}

// ------------------------------------------------------------------ (2) multiplies with ALU instructions beside them
constexpr int ILP = 8, ITERS = 4096;
__device__ __forceinline__ void unpack(uint64_t v, uint32_t &lo, uint32_t &hi) { asm("mov.b64 {%0,%1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); }

// MUL: 0 = mad.wide.u32, 1 = mad.lo.u32 ; ALU: 0 = lop3, 1 = shf, 2 = iadd3 (three register inputs) ; K ALU instructions per multiply
template <int MUL, int ALU, int K> __global__ void __launch_bounds__(256) mix(uint64_t *sink, uint32_t m0, uint32_t m1) {
    uint64_t acc[ILP];
    uint32_t z[ILP][K > 0 ? K : 1];
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        acc[i] = threadIdx.x * 977u + i;
#pragma unroll
        for (int j = 0; j < (K > 0 ? K : 1); ++j) z[i][j] = threadIdx.x + i * 7 + j;
    }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            uint32_t lo, hi; unpack(acc[i], lo, hi);
            if (MUL == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(lo), "r"(m0));
            else { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(lo) : "r"(m0), "r"(hi)); asm("mov.b64 %0, {%1,%2};" : "=l"(acc[i]) : "r"(lo), "r"(hi)); }
#pragma unroll
            for (int j = 0; j < K; ++j) {
                if (ALU == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(z[i][j]) : "r"(m0), "r"(m1));
                if (ALU == 1) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(z[i][j]) : "r"(m1));
                if (ALU == 2) asm volatile("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %0, %1;\n\taddc.u32 %0, %0, %2;\n\t}" : "+r"(z[i][j]) : "r"(m0), "r"(m1));
            }
        }
    }
    uint64_t r = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { r ^= acc[i]; for (int j = 0; j < (K > 0 ? K : 1); ++j) r ^= z[i][j]; }
    if (r == 0x12345678ull) sink[0] = r;
}
template <int MUL, int ALU, int K> void run_mix(const char *name, uint64_t *sink, int warps_per_smsp) {
    const int ctas = warps_per_smsp / 2;    // 256 threads = 2 warps per SMSP
    const int smem = ctas >= 4 ? (ctas >= 8 ? 24 * 1024 : 48 * 1024) : (ctas == 3 ? 72 * 1024 : (ctas == 2 ? 100 * 1024 : 200 * 1024));
    auto kern = mix<MUL, ALU, K>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int bps = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, 256, smem));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * bps * 2;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        kern<<<blocks, 256, smem>>>(sink, 0x9E3779B9u + rep, 0x7F4A7C15u);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const double items = (double)blocks * 256 * ITERS * ILP / (best * 1e-3);
    printf("%-44s warps/SMSP=%2d  %.2f SMSP-cycles per (1 multiply + %d ALU)\n", name, bps * 2, 148.0 * 4 * 1.965e9 / (items / 32), K);
}

// ------------------------------------------------------------------ (3) operand delivery of IMAD.WIDE
__constant__ uint32_t c_m[64];
// SRC: 0 = per-thread registers (distinct per chain), 1 = constant bank (distinct per chain), 2 = one kernel parameter (uniform register)
template <int SRC> __global__ void __launch_bounds__(256) wide_src(uint64_t *sink, uint32_t p) {
    uint64_t acc[ILP];
    uint32_t m[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { acc[i] = threadIdx.x * 977u + i; m[i] = SRC == 0 ? (threadIdx.x * 31u + i * 0x9E3779B9u) | 1u : 0u; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            uint32_t lo, hi; unpack(acc[i], lo, hi);
            const uint32_t mm = SRC == 0 ? m[i] : (SRC == 1 ? c_m[i] : p);
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(hi ^ lo), "r"(mm));
        }
    }
    uint64_t r = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) r ^= acc[i];
    if (r == 0x12345678ull) sink[0] = r;
}
template <int SRC> void run_src(const char *name, uint64_t *sink) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 8;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        wide_src<SRC><<<blocks, 256>>>(sink, 0x9E3779B9u + rep);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const double items = (double)blocks * 256 * ITERS * ILP / (best * 1e-3);
    printf("%-44s %.2f SMSP-cycles per (LOP3 + IMAD.WIDE)\n", name, 148.0 * 4 * 1.965e9 / (items / 32));
}

int main() {
    const uint32_t n = 4096;
    const uint64_t psi_inv = host::modinv(PSI, Q), omega_inv = host::mulmod(psi_inv, psi_inv, Q);
    const Mod<W> mod = host::make_mod<W>(Q, 12);
    std::vector<Tw<W>> fwd = host::fwd_pyramid<W>(PSI, n, Q), inv = host::dit_pyramid<W>(omega_inv, n, Q), last = host::fwd_last_table<W>(fwd, 12, 4);
    PolymulTables<W> tb;
    Tw<W> *d;
    CK(cudaMalloc(&d, fwd.size() * sizeof(Tw<W>))); CK(cudaMemcpy(d, fwd.data(), fwd.size() * sizeof(Tw<W>), cudaMemcpyHostToDevice)); tb.fwd_pyr = d;
    CK(cudaMalloc(&d, last.size() * sizeof(Tw<W>))); CK(cudaMemcpy(d, last.data(), last.size() * sizeof(Tw<W>), cudaMemcpyHostToDevice)); tb.fwd_last = d;
    tb.post = tb.fwd_pyr; tb.inv.pyr = tb.fwd_pyr;
    for (int i = 0; i < MAX_R; ++i) { tb.fwd_head[i] = fwd[i]; tb.inv.head[i] = inv[i]; }
    W *sink; CK(cudaMalloc(&sink, 64));
    uint32_t hm[64]; for (int i = 0; i < 64; ++i) hm[i] = 0x9E3779B9u * (i + 1) | 1u;
    CK(cudaMemcpyToSymbol(c_m, hm, sizeof(hm)));

    printf("== (1) register passes of the fused kernel, no memory traffic\n");
    run_pass<3, 0>("pass0 (uniform tw) b3", 3, tb, mod, sink);
    run_pass<3, 0>("pass0 (uniform tw) b3 @2", 2, tb, mod, sink);
    run_pass<3, 0>("pass0 (uniform tw) b3 @1", 1, tb, mod, sink);
    run_pass<2, 0>("pass0 (uniform tw) b2", 2, tb, mod, sink);
    run_pass<4, 0>("pass0 (uniform tw) b4", 4, tb, mod, sink);
    run_pass<1, 0>("pass0 (uniform tw) b1", 1, tb, mod, sink);
    run_pass<3, 1>("pass1 (broadcast tw) b3", 3, tb, mod, sink);
    run_pass<2, 1>("pass1 (broadcast tw) b2", 2, tb, mod, sink);
    run_pass<4, 1>("pass1 (broadcast tw) b4", 4, tb, mod, sink);
    run_pass<3, 2>("pass2 (per-thread tw) b3", 3, tb, mod, sink);
    run_pass<2, 2>("pass2 (per-thread tw) b2", 2, tb, mod, sink);
    run_pass<4, 2>("pass2 (per-thread tw) b4", 4, tb, mod, sink);
    run_pass<3, 3>("pass0+1+2 b3", 3, tb, mod, sink);
    run_pass<2, 3>("pass0+1+2 b2", 2, tb, mod, sink);
    run_pass<4, 3>("pass0+1+2 b4", 4, tb, mod, sink);

    printf("== (2) multiplies with ALU instructions beside them (independent chains, ILP 8)\n");
    for (int w : {16, 6}) {
        run_mix<0, 0, 0>("IMAD.WIDE alone", sink, w);
        run_mix<0, 0, 1>("IMAD.WIDE + 1 LOP3", sink, w);
        run_mix<0, 0, 2>("IMAD.WIDE + 2 LOP3", sink, w);
        run_mix<0, 0, 3>("IMAD.WIDE + 3 LOP3", sink, w);
        run_mix<0, 1, 1>("IMAD.WIDE + 1 SHF", sink, w);
        run_mix<0, 1, 2>("IMAD.WIDE + 2 SHF", sink, w);
        run_mix<0, 2, 1>("IMAD.WIDE + 1 (IADD3 + IADD3.X)", sink, w);
        run_mix<0, 2, 2>("IMAD.WIDE + 2 (IADD3 + IADD3.X)", sink, w);
        run_mix<1, 0, 0>("IMAD.LO alone", sink, w);
        run_mix<1, 0, 1>("IMAD.LO + 1 LOP3", sink, w);
        run_mix<1, 0, 2>("IMAD.LO + 2 LOP3", sink, w);
        run_mix<1, 1, 1>("IMAD.LO + 1 SHF", sink, w);
    }
    printf("== (3) multiplicand source of IMAD.WIDE (each multiply has one LOP3 beside it)\n");
    run_src<0>("registers", sink);
    run_src<1>("constant bank", sink);
    run_src<2>("kernel parameter (uniform)", sink);
    return 0;
}
