// writer micro-benchmark: pg_format_rows over 65,536 loci x 9 rows with 1..16 threads
// g++ -O2 -std=c++17 -Iinclude tools/fmt_bench.cpp -o /tmp/fmt_bench -Lpoolgen_b200/lib -lpoolgen_cuda -Wl,-rpath,$PWD/poolgen_b200/lib -lpthread
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstring>
#include "poolgen_cuda.h"
int main(int argc, char** argv) {
    const int64_t L = 65536; const int S = 3, k = 3;
    std::vector<uint64_t> meta(L), pos(L); std::vector<double> fm(L*S), st(L*S*k*4); std::vector<uint32_t> ci(L, 0);
    for (int64_t l = 0; l < L; l++) { meta[l] = 1 | (3ull << 8) | (0ull << 16) | (1ull << 24) | (2ull << 32); pos[l] = l + 1; }
    for (auto& v : fm) v = drand48(); for (auto& v : st) v = drand48() * 4 - 2;
    pg_results r{L, S, k, meta.data(), fm.data(), st.data()};
    const char* names[1] = {"chr1"};
    pg_row_labels lab{pos.data(), nullptr, nullptr, names, ci.data()};
    std::vector<char> out(L * 9 * 100); size_t n = 0;
    for (int T : {1, 2, 4, 8, 16}) {
        pg_format_rows(0, &r, &lab, T, out.data(), out.size(), &n);
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < 5; i++) pg_format_rows(0, &r, &lab, T, out.data(), out.size(), &n);
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / 5;
        printf("T=%d %.2f ms %.1f Mrows/s bytes %zu\n", T, dt * 1e3, L * 9 / dt / 1e6, n);
    }
}
