// pg_synth.h -- the synthetic workload of SURVEY.md 8(d): pure integer arithmetic shared by the device generator
// (pg_batch_synth / pg_kin_synth) and the host replay (libpoolgen_synth.so, include/poolgen_synth.h), bit for bit.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define PG_HD __host__ __device__
#else
#define PG_HD
#endif
#ifndef PG_MAX_ALLELES
#define PG_MAX_ALLELES 6
#endif

namespace pg {

// synthetic generator shared by host and device (pure integer arithmetic)
PG_HD inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// counts of one (locus, pool): c[0..A_in-1].  Locus classes (by the locus hash): 5 % monomorphic
// (fails a < 2), 2 % with pools at depth 0 (fails the depth filter), 3 % with one rare allele whose
// pooled frequency straddles 0.1 % (exercises the MAF threshold and the renormalisation of the kept
// alleles), the rest polymorphic in all of A,T,C,G.  Depth 20..100 per pool.
PG_HD inline void synth_counts(uint64_t seed, int64_t locus, int pool, int A_in, uint32_t *c) {
    const uint64_t hl = splitmix64(seed ^ ((uint64_t)locus * 0x9E3779B97F4A7C15ull));
    const uint64_t hp = splitmix64(hl ^ ((uint64_t)(pool + 1) << 20));
    const uint32_t cls = (uint32_t)(hl % 100u);
    const int n_real = A_in < 4 ? A_in : 4;  // N and D columns stay empty
    uint32_t depth = 20u + (uint32_t)(hp % 81u);
    if (cls >= 5 && cls < 7 && (uint32_t)((hl >> 32) % 61u) == (uint32_t)(pool % 61)) depth = 0;
    int rare = -1;
    uint32_t c_rare = 0;
    if (cls >= 7 && cls < 10) {
        rare = (int)((hl >> 44) % (uint64_t)n_real);
        const uint32_t thr = 30u + (uint32_t)((hl >> 50) % 93u);
        if (depth > 0 && (uint32_t)(splitmix64(hp ^ 0xBADA55ull) % 1024u) < thr) c_rare = 1;
    }
    const uint32_t rem = depth - c_rare;
    uint64_t wgt[6];
    uint64_t wsum = 0;
    for (int a = 0; a < A_in; a++) {
        const uint64_t ha = splitmix64(hl ^ (0xA11E1E00ull + (uint64_t)a));
        uint64_t base = 1 + (ha % 997u);
        if (a >= n_real || a == rare) base = 0;
        if (cls < 5) base = (a == (int)((hl >> 40) % (uint64_t)n_real)) ? 1000 : 0;  // monomorphic
        const uint64_t pa = splitmix64(hp ^ (0xC0FFEEull + (uint64_t)a));
        const uint64_t w = base * (70u + (pa % 61u));  // +-30 % per (pool, allele)
        wgt[a] = w;
        wsum += w;
    }
    uint32_t used = 0;
    int last = -1;
    for (int a = 0; a < A_in; a++) {
        const uint32_t v = wsum ? (uint32_t)(((uint64_t)rem * wgt[a]) / wsum) : 0u;
        c[a] = v;
        used += v;
        if (wgt[a]) last = a;
    }
    if (last >= 0) c[last] += rem - used;  // the remainder goes to the last allele with weight
    if (rare >= 0) c[rare] = c_rare;
}

}  // namespace pg
