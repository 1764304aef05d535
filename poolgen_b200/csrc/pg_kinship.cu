// pg_kinship.cu -- ols_iter_with_kinship (src/gwas/ols.rs:278-436): placeholder translation unit,
// the FP64 DMMA kinship GEMM and the covariate scan are built in a later milestone.
#include "pg_internal.h"
