// pg_text.cu -- sync text on the fast path: a line-aligned chunk of a sync file
// (chr \t pos \t ref \t A:T:C:G:N:D per pool ...) is copied to the device as it is and parsed there into the count slab
// the scan and the count tests consume.  Parse semantics are those of `impl Parse<LocusCounts> for String`
// (src/base/sync.rs:100-156) as driven by the chunk readers (src/base/sync.rs:827-868):
//   - a line that starts with '#' is skipped (ErrorKind::Other -> continue);
//   - a line whose second field is not an unsigned integer is skipped (same error kind);
//   - the number of pools is (tab-separated fields - 3) and must equal the scan's pool count (the reference would
//     trip the pool-size assert of the filter, src/base/sync.rs:254-257): reported as an error;
//   - every pool field holds at least six ':'-separated unsigned integers, the first six are A:T:C:G:N:D
//     (src/base/sync.rs:134-137, 144-150); anything else makes the reference panic (`expect`): reported as an error;
//   - a trailing '\r' before the newline is dropped (src/base/sync.rs:104-109).
// Pipeline: cub::DeviceSelect picks the positions of tabs + newlines and of newlines; one thread per line classifies
// it and parses the position; an exclusive scan numbers the kept lines; one warp per kept line then parses the pool
// fields (lane = pool, stride 32) straight into counts[locus][allele][pool].
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

#include "pg_device.cuh"
#include "pg_internal.h"

namespace pg {

struct IsTabOrNewline {
    const char *t;
    __host__ __device__ bool operator()(const uint32_t &i) const { return t[i] == '\t' || t[i] == '\n'; }
};
struct IsNewline {
    const char *t;
    __host__ __device__ bool operator()(const uint32_t &i) const { return t[i] == '\n'; }
};

enum { TEXT_OK = 0, TEXT_ERR_POOLS = 1, TEXT_ERR_FIELD = 2 };

struct TextParams {
    const char *text;
    uint32_t n_bytes;
    const uint32_t *sep;   // positions of '\t' and '\n', ascending
    const uint32_t *nl;    // positions of '\n', ascending
    uint32_t n_sep, n_lines;
    int n_pools;
    uint32_t *keep;        // [n_lines + 1] 1 = a locus line (then its exclusive scan)
    uint32_t *first_sep;   // [n_lines] index into sep of the first separator of the line
    uint64_t *line_pos;    // [n_lines] parsed position
    uint32_t *counts;      // [locus][6][n_pools]
    uint64_t *out_offset;  // [locus] byte offset of the line in the chunk
    uint64_t *out_pos;     // [locus]
    uint32_t *error;       // [2]: first error code, byte offset
};

__device__ __forceinline__ void text_error(const TextParams &p, uint32_t code, uint32_t at) {
    if (atomicCAS(p.error, 0u, code) == 0u) p.error[1] = at;
}

// one thread per line: comment / position / field count
__global__ void __launch_bounds__(256) text_lines_kernel(const TextParams p) {
    for (uint32_t l = blockIdx.x * blockDim.x + threadIdx.x; l < p.n_lines; l += gridDim.x * blockDim.x) {
        const uint32_t start = l ? p.nl[l - 1] + 1 : 0u, end = p.nl[l];  // [start, end) without the newline
        // rank of this line's newline among the separators (binary search), the line's separators are the ones before
        uint32_t lo = 0, hi = p.n_sep;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (p.sep[mid] < end) lo = mid + 1; else hi = mid;
        }
        const uint32_t nl_rank = lo;
        uint32_t lo2 = 0, hi2 = nl_rank;
        while (lo2 < hi2) {
            const uint32_t mid = (lo2 + hi2) >> 1;
            if (p.sep[mid] < start) lo2 = mid + 1; else hi2 = mid;
        }
        const uint32_t first = lo2;            // first tab of the line
        const uint32_t n_tabs = nl_rank - first;
        p.first_sep[l] = first;
        uint32_t keep = 0;
        if (end > start && p.text[start] != '#' && n_tabs >= 2) {
            // position = second field, `parse::<u64>()`: optional '+', then digits only
            uint32_t a = p.sep[first] + 1, b = p.sep[first + 1];
            if (a < b && p.text[a] == '+') a++;
            bool ok = a < b;
            uint64_t v = 0;
            for (uint32_t i = a; i < b && ok; i++) {
                const unsigned d = (unsigned)(p.text[i] - '0');
                if (d > 9u || v > (0xFFFFFFFFFFFFFFFFull - d) / 10ull) ok = false;
                v = v * 10ull + d;
            }
            if (ok) {
                keep = 1;
                p.line_pos[l] = v;
                if ((int)n_tabs != p.n_pools + 2) text_error(p, TEXT_ERR_POOLS, start);
            }
        }
        p.keep[l] = keep;
    }
}

// one warp per kept line: lane = pool (stride 32), six ':'-separated unsigned integers per pool field
__global__ void __launch_bounds__(256) text_parse_kernel(const TextParams p) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int n = p.n_pools;
    for (uint32_t l = warp; l < p.n_lines; l += nwarps) {
        const uint32_t ord = p.keep[l];
        if (p.keep[l + 1] == ord) continue;  // exclusive scan: not a locus line
        const uint32_t start = l ? p.nl[l - 1] + 1 : 0u;
        const uint32_t first = p.first_sep[l];
        if (lane == 0) {
            p.out_offset[ord] = start;
            p.out_pos[ord] = p.line_pos[l];
        }
        uint32_t *out = p.counts + (size_t)ord * 6 * n;
        for (int i = lane; i < n; i += 32) {
            uint32_t a = p.sep[first + 2 + i] + 1, b = p.sep[first + 3 + i];
            if (b > a && p.text[b - 1] == '\r') b--;  // Windows line end on the last pool field
            int j = 0;
            uint64_t v = 0;
            bool digits = false, bad = false;
            for (uint32_t q = a; q <= b; q++) {
                const char c = (q < b) ? p.text[q] : ':';
                if (c == ':') {
                    if (!digits) bad = true;
                    if (j < 6) out[(size_t)j * n + i] = v > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)v;
                    j++;
                    v = 0;
                    digits = false;
                } else {
                    const unsigned d = (unsigned)(c - '0');
                    if (d > 9u) bad = true;
                    v = v * 10ull + d;
                    if (v > 0xFFFFFFFFFFFFull) v = 0xFFFFFFFFFFFFull;
                    digits = true;
                }
            }
            if (bad || j < 6) text_error(p, TEXT_ERR_FIELD, a);
        }
    }
}

struct TextScratch {
    char *d_text = nullptr;
    size_t text_cap = 0;
    uint32_t *d_sep = nullptr, *d_nl = nullptr;
    size_t sep_cap = 0, nl_cap = 0;
    uint32_t *d_keep = nullptr, *d_first = nullptr;
    uint64_t *d_line_pos = nullptr;
    size_t line_cap = 0;
    uint32_t *d_num = nullptr;  // [4]: n_sep, n_lines, error code, error offset
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0;
    uint64_t *d_out_offset = nullptr, *d_out_pos = nullptr;
    uint64_t *h_out_offset = nullptr, *h_out_pos = nullptr;
    size_t out_cap = 0;
};

static cudaError_t grow(void **p, size_t *cap, size_t need, size_t elem) {
    if (*cap >= need) return cudaSuccess;
    cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    cudaError_t e = cudaMalloc(p, need * elem);
    if (e == cudaSuccess) *cap = need;
    return e;
}

void text_scratch_free(TextScratch *t) {
    if (!t) return;
    cudaFree(t->d_text);
    cudaFree(t->d_sep);
    cudaFree(t->d_nl);
    cudaFree(t->d_keep);
    cudaFree(t->d_first);
    cudaFree(t->d_line_pos);
    cudaFree(t->d_num);
    cudaFree(t->d_tmp);
    cudaFree(t->d_out_offset);
    cudaFree(t->d_out_pos);
    if (t->h_out_offset) cudaFreeHost(t->h_out_offset);
    if (t->h_out_pos) cudaFreeHost(t->h_out_pos);
    delete t;
}

// Parses `text` (host, n_bytes, ends with '\n' or at a line end) into counts[locus][6][n_pools] (device, capacity
// max_loci).  Returns the number of loci, or a negative code: -1 CUDA error (*cuda_err), -2 more loci than capacity,
// -3 pool count mismatch, -4 malformed pool field (*err_offset = byte offset in the chunk).
int64_t text_to_counts(TextScratch **scratch, const char *text, size_t n_bytes, int n_pools, uint32_t *d_counts,
                       int64_t max_loci, int sm_count, cudaStream_t s, cudaError_t *cuda_err, uint64_t *err_offset) {
#define TCK(call)                      \
    do {                               \
        cudaError_t e_ = (call);       \
        if (e_ != cudaSuccess) {       \
            *cuda_err = e_;            \
            return -1;                 \
        }                              \
    } while (0)
    if (!*scratch) *scratch = new TextScratch();
    TextScratch *t = *scratch;
    if (n_bytes == 0) return 0;
    if (n_bytes >= 0xFFFFFFF0ull) {
        *cuda_err = cudaErrorInvalidValue;
        return -1;
    }
    const bool add_nl = text[n_bytes - 1] != '\n';  // a last line without its newline
    const size_t nb = n_bytes + (add_nl ? 1 : 0);
    TCK(grow((void **)&t->d_text, &t->text_cap, nb + 16, 1));
    TCK(cudaMemcpyAsync(t->d_text, text, n_bytes, cudaMemcpyHostToDevice, s));
    if (add_nl) TCK(cudaMemsetAsync(t->d_text + n_bytes, '\n', 1, s));
    if (!t->d_num) TCK(cudaMalloc(&t->d_num, 16));
    TCK(cudaMemsetAsync(t->d_num, 0, 16, s));
    // separators: at most one per two bytes is not guaranteed, so size for the worst case lazily: first count lines
    TCK(grow((void **)&t->d_sep, &t->sep_cap, nb, 4));
    TCK(grow((void **)&t->d_nl, &t->nl_cap, nb / 2 + 16, 4));
    cub::CountingInputIterator<uint32_t> idx(0);
    size_t tmp1 = 0, tmp2 = 0, tmp3 = 0;
    cub::DeviceSelect::If(nullptr, tmp1, idx, t->d_sep, t->d_num, (int)nb, IsTabOrNewline{t->d_text}, s);
    cub::DeviceSelect::If(nullptr, tmp2, idx, t->d_nl, t->d_num + 1, (int)nb, IsNewline{t->d_text}, s);
    cub::DeviceScan::ExclusiveSum(nullptr, tmp3, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)(nb / 2 + 16), s);
    const size_t tmp = std::max(tmp1, std::max(tmp2, tmp3));
    if (t->tmp_bytes < tmp) {
        cudaFree(t->d_tmp);
        t->d_tmp = nullptr;
        t->tmp_bytes = 0;
        TCK(cudaMalloc(&t->d_tmp, tmp));
        t->tmp_bytes = tmp;
    }
    size_t tb = t->tmp_bytes;
    TCK(cub::DeviceSelect::If(t->d_tmp, tb, idx, t->d_sep, t->d_num, (int)nb, IsTabOrNewline{t->d_text}, s));
    tb = t->tmp_bytes;
    TCK(cub::DeviceSelect::If(t->d_tmp, tb, idx, t->d_nl, t->d_num + 1, (int)nb, IsNewline{t->d_text}, s));
    uint32_t num[2];
    TCK(cudaMemcpyAsync(num, t->d_num, 8, cudaMemcpyDeviceToHost, s));
    TCK(cudaStreamSynchronize(s));
    const uint32_t n_sep = num[0], n_lines = num[1];
    if (n_lines == 0) return 0;
    if (t->line_cap < (size_t)n_lines + 1) {
        cudaFree(t->d_keep);
        cudaFree(t->d_first);
        cudaFree(t->d_line_pos);
        t->d_keep = nullptr, t->d_first = nullptr, t->d_line_pos = nullptr;
        t->line_cap = 0;
        TCK(cudaMalloc(&t->d_keep, ((size_t)n_lines + 1) * 4));
        TCK(cudaMalloc(&t->d_first, (size_t)n_lines * 4));
        TCK(cudaMalloc(&t->d_line_pos, (size_t)n_lines * 8));
        t->line_cap = (size_t)n_lines + 1;
    }
    if (t->out_cap < (size_t)max_loci) {
        cudaFree(t->d_out_offset);
        cudaFree(t->d_out_pos);
        if (t->h_out_offset) cudaFreeHost(t->h_out_offset);
        if (t->h_out_pos) cudaFreeHost(t->h_out_pos);
        t->d_out_offset = t->d_out_pos = nullptr;
        t->h_out_offset = t->h_out_pos = nullptr;
        t->out_cap = 0;
        TCK(cudaMalloc(&t->d_out_offset, (size_t)max_loci * 8));
        TCK(cudaMalloc(&t->d_out_pos, (size_t)max_loci * 8));
        TCK(cudaHostAlloc((void **)&t->h_out_offset, (size_t)max_loci * 8, cudaHostAllocDefault));
        TCK(cudaHostAlloc((void **)&t->h_out_pos, (size_t)max_loci * 8, cudaHostAllocDefault));
        t->out_cap = (size_t)max_loci;
    }
    TextParams tp;
    tp.text = t->d_text;
    tp.n_bytes = (uint32_t)nb;
    tp.sep = t->d_sep;
    tp.nl = t->d_nl;
    tp.n_sep = n_sep;
    tp.n_lines = n_lines;
    tp.n_pools = n_pools;
    tp.keep = t->d_keep;
    tp.first_sep = t->d_first;
    tp.line_pos = t->d_line_pos;
    tp.counts = d_counts;
    tp.out_offset = t->d_out_offset;
    tp.out_pos = t->d_out_pos;
    tp.error = t->d_num + 2;
    const int g1 = (int)std::min<uint32_t>((n_lines + 255) / 256, (uint32_t)sm_count * 8);
    text_lines_kernel<<<g1, 256, 0, s>>>(tp);
    TCK(cudaGetLastError());
    TCK(cudaMemsetAsync(t->d_keep + n_lines, 0, 4, s));
    tb = t->tmp_bytes;
    TCK(cub::DeviceScan::ExclusiveSum(t->d_tmp, tb, t->d_keep, t->d_keep, (int)(n_lines + 1), s));
    uint32_t n_loci = 0, err[2] = {0, 0};
    TCK(cudaMemcpyAsync(&n_loci, t->d_keep + n_lines, 4, cudaMemcpyDeviceToHost, s));
    TCK(cudaMemcpyAsync(err, t->d_num + 2, 8, cudaMemcpyDeviceToHost, s));
    TCK(cudaStreamSynchronize(s));
    if (err[0]) {
        *err_offset = err[1];
        return err[0] == TEXT_ERR_POOLS ? -3 : -4;
    }
    if ((int64_t)n_loci > max_loci) return -2;
    if (n_loci == 0) return 0;
    const int g2 = (int)std::min<uint64_t>(((uint64_t)n_lines * 32 + 255) / 256, (uint64_t)sm_count * 8);
    text_parse_kernel<<<g2, 256, 0, s>>>(tp);
    TCK(cudaGetLastError());
    TCK(cudaMemcpyAsync(err, t->d_num + 2, 8, cudaMemcpyDeviceToHost, s));
    TCK(cudaMemcpyAsync(t->h_out_offset, t->d_out_offset, (size_t)n_loci * 8, cudaMemcpyDeviceToHost, s));
    TCK(cudaMemcpyAsync(t->h_out_pos, t->d_out_pos, (size_t)n_loci * 8, cudaMemcpyDeviceToHost, s));
    TCK(cudaStreamSynchronize(s));
    if (err[0]) {
        *err_offset = err[1];
        return err[0] == TEXT_ERR_POOLS ? -3 : -4;
    }
    return (int64_t)n_loci;
#undef TCK
}

const uint64_t *text_offsets(const TextScratch *t) { return t ? t->h_out_offset : nullptr; }
const uint64_t *text_positions(const TextScratch *t) { return t ? t->h_out_pos : nullptr; }

}  // namespace pg
