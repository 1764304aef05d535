/*
 * poolgen_synth.h -- host replay of the synthetic workload, libpoolgen_synth.so.
 *
 * Plain C, no CUDA: the same integer-hash generator pg_batch_synth / pg_kin_synth run on the device
 * (poolgen_b200/csrc/pg_synth.h), so the CPU checker and the CPU baseline see the bits the GPU sees without mapping
 * libpoolgen_cuda.so.  Returns 0 on success, -1 on a bad argument.
 */
#ifndef POOLGEN_SYNTH_H
#define POOLGEN_SYNTH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* counts u32 [locus][allele][pool] (pool fastest), identical bits to pg_batch_synth */
int pg_synth_counts_host(uint64_t seed, int64_t first_locus, int64_t n_loci, int n_pools, int n_alleles,
                         uint32_t *counts_out);
int pg_synth_phen_host(uint64_t seed, int n_pools, int k, double *phen_out /* n_pools x k row-major */);
/* the same counts as sync text (six columns, N = D = 0 beyond n_alleles): chr<1 + locus / 1000000> \t <locus + 1> \t N ...;
 * returns the bytes written (or needed, when capacity is too small) in *n_bytes */
int pg_synth_sync_text_host(uint64_t seed, int64_t first_locus, int64_t n_loci, int n_pools, int n_alleles,
                            char *out, size_t capacity, size_t *n_bytes);

#ifdef __cplusplus
}
#endif
#endif
