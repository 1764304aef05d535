// pg_kin.h -- the handle of the ols_iter_with_kinship path (shared by pg_kinship.cu and pg_comm.cu; not part of the ABI)
#pragma once
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "pg_internal.h"

struct pg_kin {
    pg_ctx *ctx = nullptr;
    int n = 0, ldg = 0;
    int64_t cap = 0, P = 0;
    int64_t P_total = 0;      // columns over all shards (pg_kin_allreduce); 0 = this handle holds them all
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double *d_G = nullptr;
    double *d_K = nullptr;    // [n][n] partial Gram matrix (sum over the resident columns)
    double *d_ws = nullptr;   // split-K workspace
    size_t ws_bytes = 0;
    int nt = 0, n_slices = 0;
    // covariates
    int m = -1;
    bool minnorm = false;         // m == n: the reference's n < p branch (pg_kin_eig_select)
    std::vector<double> eigvals;  // descending
    std::vector<double> Q;        // [1+m][ldg] orthonormal basis of [1 | PCs] (host)
    std::vector<double> C;        // [m][ldg] the covariates themselves (pg_kin_mle_scan: the simplex search is not basis-invariant)
    double *d_mle = nullptr;      // pg_kin_mle_scan: centred covariates and phenotypes, then the fixed moments
    size_t mle_bytes = 0;
    double *d_V = nullptr;        // [(1+m) + k][ldg]
    size_t V_bytes = 0;
    double *d_ptab = nullptr;
    double ptab_isd = 0, ptab_bits = 0;
    int ptab_M = 0;
    // results
    int k = 0;
    double *d_res = nullptr;  // [3][k][P]
    double *h_res = nullptr;  // pinned
    size_t res_elems = 0;
    void *d_defer = nullptr;  // covariate scan: [count | columns left to the two-pass kernel]
    size_t defer_bytes = 0;
    double *d_part = nullptr;  // covariate scan (DMMA form in pool passes): partial sums per column block
    size_t part_bytes = 0;
    // loader scratch
    uint32_t *d_sel = nullptr;
    int64_t *d_off = nullptr;
    int64_t *d_col_locus = nullptr;
    uint8_t *d_col_allele = nullptr;
    void *d_scan_tmp = nullptr;
    size_t scan_tmp_bytes = 0;
    int64_t load_cap = 0;
    void *d_counts = nullptr;
    size_t counts_bytes = 0;
    double *d_w = nullptr;
    pg::TextScratch *text = nullptr;  // sync text parsed on the device (pg_kin_append_sync_text)
    // eigen step: cuSOLVER is mapped and its handle created by a background thread started in pg_kin_open and joined
    // by pg_kin_eig_select / pg_kin_close
    std::thread warm;
    void *solver = nullptr;  // cusolverDnHandle_t
};

