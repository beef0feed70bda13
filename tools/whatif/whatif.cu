// Stand-alone timing harness for ONE instantiation of the fused polymul kernel (north-star parameter set:
// N = 4096, q = 2^60 - 2^14 + 1).  It exists so that a scheduling experiment (a -D toggle in kernels.cuh /
// modarith.cuh, or different template parameters) can be compiled in seconds and timed in one short GPU
// call without rebuilding libtntt.so.  Row 0 of the result is checked against a schoolbook negacyclic
// product computed on the host with 128-bit arithmetic (unless -DWHATIF_WRONG_RESULTS, for "what if this
// cost nothing" experiments that knowingly break the arithmetic).
//
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo [-D...] -o whatif_X whatif.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>

#include "../../tiny-ntt_b200/csrc/common.h"
#include "../../tiny-ntt_b200/csrc/tables.h"

#ifndef W_LOGR
#define W_LOGR 4
#endif
#ifndef W_NA
#define W_NA 1
#endif
#ifndef W_MINB
#define W_MINB 3
#endif
#ifndef W_STASH
#define W_STASH 1
#endif
#ifndef W_TMA
#define W_TMA 0
#endif
#ifndef W_ROWS
#define W_ROWS 16384
#endif
#ifndef W_NAME
#define W_NAME "baseline"
#endif

using namespace tntt;
#ifdef W_U32   // the 24-bit N = 4096 set of the reference (32-bit words, no intermediate reduction)
using W = uint32_t;
constexpr uint64_t Q = 8380417ull, PSI = 283817ull;
constexpr bool kRed = false;
#else
using W = uint64_t;
constexpr uint64_t Q = 1152921504606830593ull, PSI = 431606828070683274ull;
constexpr bool kRed = true;
#endif
#ifndef W_PAD
#define W_PAD 0
#endif
#ifndef W_RED
#define W_RED (kRed ? 1 : 0)     // 2 = Solinas reductions (q = 2^60 - 2^14 + 1 only)
#endif
using C = Cfg<W, 12, W_LOGR, 1, W_PAD>;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

template <typename T> T *upload(const std::vector<T> &h) {
    T *d; CK(cudaMalloc(&d, h.size() * sizeof(T)));
    CK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return d;
}

int main(int argc, char **argv) {
    const size_t rows = argc > 1 ? (size_t)atol(argv[1]) : W_ROWS;
    const uint32_t n = 4096;
    const uint64_t psi_inv = host::modinv(PSI, Q), omega_inv = host::mulmod(psi_inv, psi_inv, Q), n_inv = host::modinv(n, Q);
    const Mod<W> mod = host::make_mod<W>(Q, 12);
    std::vector<Tw<W>> fwd = host::fwd_pyramid<W>(PSI, n, Q), inv = host::dit_pyramid<W>(omega_inv, n, Q);
    const uint64_t r_mod_q = (uint64_t)((((host::u128)1) << (8 * sizeof(W))) % Q);
    PolymulTables<W> tb;
    tb.fwd_pyr = upload(fwd);
    tb.fwd_last = upload(host::fwd_last_table<W>(fwd, 12, W_LOGR));
    tb.post = upload(host::scaled_powers<W>(psi_inv, W_RED == 2 ? n_inv : host::mulmod(n_inv, r_mod_q, Q), n, Q));   // Solinas pointwise: no 2^-64
    tb.inv.pyr = upload(inv);
    for (int i = 0; i < MAX_R; ++i) { tb.fwd_head[i] = fwd[i]; tb.inv.head[i] = inv[i]; }

    std::vector<W> a(rows * n), b(rows * n);
    uint64_t s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s % Q; };
    for (size_t i = 0; i < rows * n; ++i) { a[i] = (W)rnd(); b[i] = (W)rnd(); }
    for (uint32_t i = 0; i < n; ++i) { a[n + i] = (W)(Q - 1); b[n + i] = (W)(Q - 1); }   // row 1: worst case for the lazy ranges
    W *da = upload(a), *db = upload(b), *dc;
    CK(cudaMalloc(&dc, rows * n * sizeof(W)));

#if defined(W_CLUSTER)
    auto kern = polymul_cluster_kernel<C, W_CLUSTER, W_RED>;
    const size_t smem = 2 * (C::N / W_CLUSTER) * sizeof(W);
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaFuncAttributes attr; CK(cudaFuncGetAttributes(&attr, kern));
    int bps = 0;
    auto launch = [&]() {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(rows * W_CLUSTER)); cfg.blockDim = dim3(C::P / W_CLUSTER); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = W_CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, kern, (const W *)da, (const W *)db, dc, rows, tb, mod));
    };
#else
    auto kern = polymul_kernel<C, W_NA, W_RED, W_MINB, W_STASH, W_TMA>;
    const size_t smem = ((size_t)W_NA * C::TILE + (W_STASH == 1 ? C::N : 0)) * sizeof(W) + (W_TMA ? kTwBufBytes + 16 : 0);
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaFuncAttributes attr; CK(cudaFuncGetAttributes(&attr, kern));
    int bps = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, C::THREADS, smem));
#ifdef TNTT_X_PERSISTENT
    const unsigned grid = (unsigned)std::min<size_t>(rows, (size_t)148 * bps);
#else
    const unsigned grid = (unsigned)rows;
#endif
    auto launch = [&]() { kern<<<grid, C::THREADS, smem>>>(da, db, dc, rows, tb, mod); };
#endif
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f, sum = 0;
    const int reps = 5, per = rows <= 1024 ? 50 : 4;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        for (int i = 0; i < per; ++i) launch();
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= per;
        sum += ms; if (ms < best) best = ms;
    }
    std::vector<W> c(2 * n);
    CK(cudaMemcpy(c.data(), dc, 2 * n * sizeof(W), cudaMemcpyDeviceToHost));
    const char *verdict = "unchecked";
#ifndef WHATIF_WRONG_RESULTS
    long bad = 0;
    for (int row = 0; row < 2; ++row) {
        std::vector<uint64_t> want(n, 0);
        for (uint32_t i = 0; i < n; ++i)
            for (uint32_t j = 0; j < n; ++j) {
                const uint64_t p = host::mulmod(a[row * n + i], b[row * n + j], Q);
                const uint32_t k = (i + j) & (n - 1);
                if (i + j < n) { want[k] += p; if (want[k] >= Q) want[k] -= Q; }
                else { want[k] = want[k] >= p ? want[k] - p : want[k] + Q - p; }
            }
        for (uint32_t k = 0; k < n; ++k) bad += want[k] != c[row * n + k];
    }
    verdict = bad ? "WRONG" : "exact";
#endif
    printf("%-28s us/launch=%.2f rows=%zu regs=%d local=%zu smem=%zu ctas/SM=%d  mean %.4f ms  best %.4f ms  %.3f Mpolymul/s (mean) %.3f (best)  rows0-1 %s\n",
           W_NAME, best * 1e3, rows, attr.numRegs, (size_t)attr.localSizeBytes, smem, bps, sum / reps, best, rows / (sum / reps) / 1e3, rows / best / 1e3, verdict);
    return 0;
}
