// pg_scan_inst.cu -- one translation unit per device allele count (compiled with -DPG_SCAN_A=2..6) so the
// template instantiations of pg_scan.cuh build in parallel.
#include "pg_scan.cuh"

#ifndef PG_SCAN_A
#error "compile with -DPG_SCAN_A=<2..6>"
#endif

namespace pg {
#define PG_CAT2(a, b) a##b
#define PG_CAT(a, b) PG_CAT2(a, b)
cudaError_t PG_CAT(launch_scan_a, PG_SCAN_A)(const ScanParams &p, int sm_count, cudaStream_t s) {
    return launch_scan_a<PG_SCAN_A>(p, sm_count, s);
}
}  // namespace pg
