// pg_synth_host.cpp -- host replay of the synthetic workload (include/poolgen_synth.h, libpoolgen_synth.so).
// A library of its own, without CUDA: bench.py's reference arm and the CPU tests generate their inputs here and never
// map the product library.
#include <stdio.h>
#include <string.h>

#include "pg_synth.h"
#include "poolgen_synth.h"

extern "C" {

// ---- synthetic workload, host replay -----------------------------------------------------------
int pg_synth_counts_host(uint64_t seed, int64_t first_locus, int64_t n_loci, int n_pools, int n_alleles,
                         uint32_t *out) {
    if (!out || n_pools < 1 || n_alleles < 1 || n_alleles > PG_MAX_ALLELES || n_loci < 0) return -1;
    for (int64_t l = 0; l < n_loci; l++)
        for (int i = 0; i < n_pools; i++) {
            uint32_t c[PG_MAX_ALLELES];
            pg::synth_counts(seed, first_locus + l, i, n_alleles, c);
            for (int a = 0; a < n_alleles; a++) out[((size_t)l * n_alleles + a) * n_pools + i] = c[a];
        }
    return 0;
}

int pg_synth_sync_text_host(uint64_t seed, int64_t first_locus, int64_t n_loci, int n_pools, int n_alleles, char *out,
                            size_t capacity, size_t *n_bytes) {
    if (!n_bytes || n_pools < 1 || n_alleles < 1 || n_alleles > PG_MAX_ALLELES || n_loci < 0) return -1;
    size_t w = 0;
    char tmp[64];
    auto put = [&](const char *p, size_t len) {
        if (out && w + len <= capacity) memcpy(out + w, p, len);
        w += len;
    };
    for (int64_t l = 0; l < n_loci; l++) {
        const int64_t locus = first_locus + l;
        int len = snprintf(tmp, sizeof tmp, "chr%lld\t%lld\tN", (long long)(1 + locus / 1000000), (long long)(locus + 1));
        put(tmp, (size_t)len);
        for (int i = 0; i < n_pools; i++) {
            uint32_t c[PG_MAX_ALLELES] = {0, 0, 0, 0, 0, 0};
            pg::synth_counts(seed, locus, i, n_alleles, c);
            len = snprintf(tmp, sizeof tmp, "\t%u:%u:%u:%u:%u:%u", c[0], c[1], c[2], c[3], c[4], c[5]);
            put(tmp, (size_t)len);
        }
        put("\n", 1);
    }
    *n_bytes = w;
    return (out && w <= capacity) ? 0 : -1;
}

int pg_synth_phen_host(uint64_t seed, int n_pools, int k, double *out) {
    if (!out || n_pools < 1 || k < 1) return -1;
    for (int i = 0; i < n_pools; i++)
        for (int j = 0; j < k; j++) {
            const uint64_t h = pg::splitmix64(seed ^ 0x9E11E5ull ^ ((uint64_t)(i + 1) << 24) ^ (uint64_t)j);
            // 53 uniform bits -> [-3, 3)
            out[(size_t)i * k + j] = ((double)(h >> 11) * (1.0 / 9007199254740992.0)) * 6.0 - 3.0;
        }
    return 0;
}

}  // extern "C"
