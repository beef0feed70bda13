// Thin C-ABI shim around the UNMODIFIED reference benchmark sources.
// TEST INFRASTRUCTURE ONLY (see oracle/ntt_oracle.py header).
//
// The reference's software_benchmark/benchmark_ntt.cpp and
// benchmark_ntt_60bit.cpp keep everything in an anonymous namespace next to
// main(), so they cannot be linked against.  This file #includes the source
// where it lies under /root/reference (BENCH_SRC, passed by oracle/Makefile;
// nothing is copied into this repository), renames its main(), and exports
// the reference's own functions with C linkage so tests and bench.py can run
// them on arbitrary inputs:
//   negacyclic_mul_ntt        benchmark_ntt.cpp:194-205 / benchmark_ntt_60bit.cpp:148-159
//   forward_ntt_bench         benchmark_ntt.cpp:207-211 / _60bit.cpp:161-165
//   negacyclic_mul_reference  benchmark_ntt.cpp:213-226 / _60bit.cpp:167-180
//   make_poly, checksum       benchmark_ntt.cpp:82-90,228-233 / _60bit.cpp:79-87,182-188
// BENCH_N / BENCH_Q / BENCH_PSI / BENCH_SIMD_KIND are the reference's own
// compile-time parameters (software_benchmark/CMakeLists.txt:20-27).
#define main tntt_ref_original_main
#include BENCH_SRC
#undef main

#include <cstring>
#include <thread>
#include <vector>

namespace {
using Word = Poly::value_type;

void run_rows(const Word *a, const Word *b, Word *c, std::size_t lo, std::size_t hi) {
    Poly pa, pb, pc;
    for (std::size_t r = lo; r < hi; ++r) {
        std::memcpy(pa.data(), a + r * N, N * sizeof(Word));
        std::memcpy(pb.data(), b + r * N, N * sizeof(Word));
        negacyclic_mul_ntt(pa, pb, pc);
        std::memcpy(c + r * N, pc.data(), N * sizeof(Word));
    }
}
}  // namespace

extern "C" {

unsigned long long tntt_ref_n() { return N; }
unsigned long long tntt_ref_q() { return Q; }
unsigned long long tntt_ref_psi() { return PSI; }
int tntt_ref_word_bytes() { return (int)sizeof(Word); }
int tntt_ref_simd_kind() { return BENCH_SIMD_KIND; }

// rows of a [batch, N] array of Word, `threads` host threads, one reference call per row
void tntt_ref_polymul_batch(const void *a, const void *b, void *c, unsigned long long batch, int threads) {
    const Word *pa = static_cast<const Word *>(a), *pb = static_cast<const Word *>(b);
    Word *pc = static_cast<Word *>(c);
    if (threads < 1) threads = 1;
    if ((unsigned long long)threads > batch) threads = batch ? (int)batch : 1;
    if (threads == 1) { run_rows(pa, pb, pc, 0, batch); return; }
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back(run_rows, pa, pb, pc, batch * t / threads, batch * (t + 1) / threads);
    for (auto &th : pool) th.join();
}

void tntt_ref_forward(const void *a, void *out) {
    Poly pa, po;
    std::memcpy(pa.data(), a, N * sizeof(Word));
    forward_ntt_bench(pa, po);
    std::memcpy(out, po.data(), N * sizeof(Word));
}

void tntt_ref_schoolbook(const void *a, const void *b, void *c) {
    Poly pa, pb, pc;
    std::memcpy(pa.data(), a, N * sizeof(Word));
    std::memcpy(pb.data(), b, N * sizeof(Word));
    negacyclic_mul_reference(pa, pb, pc);
    std::memcpy(c, pc.data(), N * sizeof(Word));
}

void tntt_ref_make_poly(unsigned long long seed, void *out) {
    const Poly p = make_poly(static_cast<decltype(make_poly(0))::value_type>(seed));
    std::memcpy(out, p.data(), N * sizeof(Word));
}

unsigned long long tntt_ref_checksum(const void *v) {
    Poly p;
    std::memcpy(p.data(), v, N * sizeof(Word));
    return checksum(p);
}

}  // extern "C"
