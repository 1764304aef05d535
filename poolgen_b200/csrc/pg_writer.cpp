// pg_writer.cpp -- the CSV rows of the reference's writers from the numeric records of a scan.
// Host code (the text ends up in a host file either way; records are smaller than rows, so they cross PCIe):
//   ols_iterate     src/gwas/ols.rs:255-275            chr,pos,allele,freq r8,Pheno_j,beta r6,p r12
//   correlation     src/gwas/correlation_test.rs:113-128  chr,pos,allele,freq,Pheno_j,r r6,p
//   chisq           src/tables/chisq_test.rs:36-46     chr,pos,alleles,chi2 r6,p
//   fisher          src/tables/fisher_exact_test.rs:118-129  chr,pos,alleles,p_observed,p
//   ols_with_covariate  src/gwas/ols.rs:409-433        chr,pos,allele,Pheno_j,beta,p  (phenotype outer, column inner)
//   sync2csv        src/base/sync.rs:1182-1262         chr,pos,allele,f_pool1 r6,...  (loci sorted by chromosome, position)
// Numbers follow Rust's `f64::to_string()` (shortest digits that round-trip, never an exponent, "NaN", "inf", "-0")
// and src/base/helpers.rs:103-117 (`sensible_round`, `parse_f64_roundup_and_own`).  Loci are split over threads in
// contiguous ranges; the pieces are concatenated in locus order, like the reference concatenates its chunk files
// (src/base/sync.rs:953-967).
#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <cfloat>

#include "pg_statrs_host.h"
#include "poolgen_cuda.h"

namespace {

const char kAlleleNames[6] = {'A', 'T', 'C', 'G', 'N', 'D'};  // src/base/sync.rs:134-137

inline char *put_u64(uint64_t v, char *o) {
    auto r = std::to_chars(o, o + 24, v);
    return r.ptr;
}

// digits d[0..nd) with the decimal point after d[0] scaled by 10^e10 -> positional notation
inline char *put_positional(bool neg, const char *d, int nd, int e10, char *o) {
    if (neg) *o++ = '-';
    if (e10 >= nd - 1) {
        memcpy(o, d, (size_t)nd);
        o += nd;
        const int z = e10 - (nd - 1);
        memset(o, '0', (size_t)z);
        o += z;
    } else if (e10 >= 0) {
        memcpy(o, d, (size_t)e10 + 1);
        o += e10 + 1;
        *o++ = '.';
        memcpy(o, d + e10 + 1, (size_t)(nd - e10 - 1));
        o += nd - e10 - 1;
    } else {
        *o++ = '0';
        *o++ = '.';
        memset(o, '0', (size_t)(-e10 - 1));
        o += -e10 - 1;
        memcpy(o, d, (size_t)nd);
        o += nd;
    }
    return o;
}

// Rust `Display for f64`: at most 1 + 1 + 324 + 17 characters
char *put_f64(double x, char *o) {
    if (x != x) {
        memcpy(o, "NaN", 3);
        return o + 3;
    }
    if (std::isinf(x)) {
        if (x < 0) *o++ = '-';
        memcpy(o, "inf", 3);
        return o + 3;
    }
    if (x == 0.0) {
        if (std::signbit(x)) *o++ = '-';
        *o++ = '0';
        return o;
    }
    char tmp[40];
    auto r = std::to_chars(tmp, tmp + sizeof tmp, x, std::chars_format::scientific);  // shortest round-trip digits
    const char *p = tmp;
    bool neg = false;
    if (*p == '-') {
        neg = true;
        p++;
    }
    char d[24];
    int nd = 0;
    for (; p < r.ptr && *p != 'e'; p++)
        if (*p != '.') d[nd++] = *p;
    p++;  // 'e'
    bool eneg = false;
    if (*p == '-') {
        eneg = true;
        p++;
    } else if (*p == '+') {
        p++;
    }
    int e10 = 0;
    for (; p < r.ptr; p++) e10 = e10 * 10 + (*p - '0');
    if (eneg) e10 = -e10;
    while (nd > 1 && d[nd - 1] == '0') nd--;
    return put_positional(neg, d, nd, e10, o);
}

struct Rounder {
    int digits;
    double factor;  // the correctly rounded parse of "1e<digits>" (helpers.rs:104-106)
    explicit Rounder(int nd) : digits(nd) {
        char f[16];
        snprintf(f, sizeof f, "1e%d", nd);
        factor = strtod(f, nullptr);
    }
    // parse_f64_roundup_and_own (helpers.rs:111-117)
    char *put(double x, char *o) const {
        const double y = x * factor;
        if (std::fabs(y) < 9.0e14) {
            // k = round(y) is an integer below 10^15: the shortest digits of k / factor ARE the decimal k / 10^digits
            // (two different decimals of <= 15 significant digits never share a double), and an x whose own shortest
            // form is shorter than `digits` characters has fewer than `digits` decimals, so rounding returns it
            // unchanged -- the `s.len() < n_digits` shortcut of the reference cannot be observed in this range
            const double kd = std::round(y);
            if (kd == 0.0) {
                // -0 survives the division (Rust prints "-0")
                if (std::signbit(kd)) *o++ = '-';
                *o++ = '0';
                return o;
            }
            const bool neg = kd < 0;
            uint64_t k = (uint64_t)(neg ? -kd : kd);
            char d[24];
            int nd = 0;
            {
                char rev[24];
                int n = 0;
                while (k) {
                    rev[n++] = (char)('0' + k % 10);
                    k /= 10;
                }
                while (n) d[nd++] = rev[--n];
            }
            const int e10 = nd - 1 - digits;
            while (nd > 1 && d[nd - 1] == '0') nd--;
            return put_positional(neg, d, nd, e10, o);
        }
        char tmp[400];
        char *e = put_f64(x, tmp);
        if ((int)(e - tmp) < digits) {
            memcpy(o, tmp, (size_t)(e - tmp));
            return o + (e - tmp);
        }
        return put_f64(std::round(y) / factor, o);
    }
};

struct Labels {
    const pg_row_labels *lab;
    // chromosome text of locus l
    inline void chr(int64_t l, const char *&s, size_t &n) const {
        if (lab->text) {
            s = lab->text + lab->line_offsets[l];
            const char *t = s;
            while (*t != '\t') t++;
            n = (size_t)(t - s);
        } else {
            s = lab->chr_names[lab->chr_index[l]];
            n = strlen(s);
        }
    }
};

void format_range(int kind, const pg_results *res, const pg_row_labels *lab, int64_t lo, int64_t hi, std::string &dst,
                  int flags = 0, int n_pools = 0) {
    // rows are appended to a string that lives on this thread's stack and handed over once: the std::string headers of
    // the per-thread pieces sit side by side in one vector, and appending through them would bounce their cache line
    // between the cores on every row
    std::string out;
    struct Handover {
        std::string &from, &to;
        ~Handover() { to = std::move(from); }
    } handover{out, dst};
    const Labels L{lab};
    const int S = res->n_slots, k = res->n_phen;
    const Rounder r6(6), r8(8), r12(12);
    // PG_FORMAT_EXACT_P: the p-value is re-derived on the host from the record's t statistic with the reference's own
    // arithmetic (statrs' continued fraction, df = n - 1 for ols_iter, n - 2 for pearson_corr)
    const bool exact_p = (flags & PG_FORMAT_EXACT_P) && (kind == PG_KIND_OLS || kind == PG_KIND_CORR || kind == PG_KIND_MLE) &&
                         (double)n_pools - (kind != PG_KIND_CORR ? 1.0 : 2.0) > 0.0;
    const pg::statrs::TwoSidedT tail(exact_p ? (double)n_pools - (kind != PG_KIND_CORR ? 1.0 : 2.0) : 1.0);
    char line[1024];
    out.reserve((size_t)(hi - lo) * 64 * (size_t)(S * k > 0 ? S * k : 1) / 2 + 4096);
    for (int64_t l = lo; l < hi; l++) {
        const uint64_t mv = res->meta[l];
        if ((mv & 0xffu) != PG_LOCUS_OK) continue;  // None: no row (src/base/sync.rs:864-867)
        const int n_out = (int)((mv >> 8) & 0xffu);
        const char *cs;
        size_t cn;
        L.chr(l, cs, cn);
        char head[320];
        char *h = head;
        if (cn > 256) cn = 256;
        memcpy(h, cs, cn);
        h += cn;
        *h++ = ',';
        h = put_u64(lab->positions[l], h);
        *h++ = ',';
        const size_t hn = (size_t)(h - head);
        if (kind == PG_KIND_GWALPHA_LS || kind == PG_KIND_GWALPHA_ML) {
            // gwalpha_ls / gwalpha_ml (src/gwas/gwalpha.rs:320-331): chr,pos,allele,freq r6,Pheno_0,alpha r6,Unknown
            for (int s = 0; s < n_out && s < S; s++) {
                const double *st = res->stats + ((size_t)l * S + s) * (size_t)k * 4;
                char *o = line;
                memcpy(o, head, hn);
                o += hn;
                *o++ = kAlleleNames[(mv >> (16 + 8 * s)) & 0xffu];
                *o++ = ',';
                o = r6.put(res->freq_mean[(size_t)l * S + s], o);
                memcpy(o, ",Pheno_0,", 9);
                o += 9;
                o = r6.put(st[0], o);
                memcpy(o, ",Unknown\n", 9);
                o += 9;
                out.append(line, (size_t)(o - line));
            }
        } else if (kind == PG_KIND_OLS || kind == PG_KIND_CORR || kind == PG_KIND_MLE) {
            for (int s = 0; s < n_out && s < S; s++) {
                const double fm = res->freq_mean[(size_t)l * S + s];
                const char an = kAlleleNames[(mv >> (16 + 8 * s)) & 0xffu];
                for (int j = 0; j < k; j++) {
                    const double *st = res->stats + (((size_t)l * S + s) * k + j) * 4;
                    char *o = line;
                    memcpy(o, head, hn);
                    o += hn;
                    *o++ = an;
                    *o++ = ',';
                    o = (kind != PG_KIND_CORR) ? r8.put(fm, o) : put_f64(fm, o);
                    memcpy(o, ",Pheno_", 7);
                    o += 7;
                    o = put_u64((uint64_t)j, o);
                    *o++ = ',';
                    o = r6.put(st[0], o);
                    *o++ = ',';
                    double pv = st[3];
                    // a finite or infinite t: the reference's formula (ols.rs:148-154, correlation_test.rs:62-66);
                    // a NaN t marks the special cases the record already carries (p forced to 1, eps or NaN)
                    if (exact_p && st[2] == st[2] && fabs(st[2]) > DBL_EPSILON) pv = tail(fabs(st[2]));
                    o = (kind == PG_KIND_OLS) ? r12.put(pv, o) : put_f64(pv, o);
                    *o++ = '\n';
                    out.append(line, (size_t)(o - line));
                }
            }
        } else {
            const double *st = res->stats + (size_t)l * 4;
            char *o = line;
            memcpy(o, head, hn);
            o += hn;
            for (int s = 0; s < n_out && s < PG_MAX_ALLELES; s++) *o++ = kAlleleNames[(mv >> (16 + 8 * s)) & 0xffu];
            *o++ = ',';
            o = (kind == PG_KIND_CHISQ) ? r6.put(st[0], o) : put_f64(st[0], o);
            *o++ = ',';
            o = put_f64(st[3], o);
            *o++ = '\n';
            out.append(line, (size_t)(o - line));
        }
    }
}

int emit(const std::vector<std::string> &parts, char *out, size_t capacity, size_t *n_bytes) {
    size_t total = 0;
    for (const auto &p : parts) total += p.size();
    if (n_bytes) *n_bytes = total;
    if (total > capacity || (!out && total)) return PG_ERR_ARG;
    size_t o = 0;
    for (const auto &p : parts) {
        memcpy(out + o, p.data(), p.size());
        o += p.size();
    }
    return PG_OK;
}

}  // namespace

extern "C" {

int pg_format_header(int kind, char *out, size_t capacity, size_t *n_bytes) {
    const char *h;
    switch (kind) {
        case PG_KIND_OLS:
        case PG_KIND_MLE:
        case PG_KIND_GWALPHA_LS:
        case PG_KIND_GWALPHA_ML:
        case PG_KIND_CORR: h = "#chr,pos,alleles,freq,phenotype,statistic,pvalue\n"; break;  // src/base/sync.rs:950
        case PG_KIND_CHISQ:
        case PG_KIND_FISHER: h = "#chr,pos,alleles,statistic,pvalue\n"; break;               // src/base/sync.rs:766
        case PG_KIND_OLS_KINSHIP: h = "#chr,pos,alleles,phenotype,statistic,pvalue\n"; break;  // src/gwas/ols.rs:409
        default: return PG_ERR_ARG;
    }
    const size_t n = strlen(h);
    if (n_bytes) *n_bytes = n;
    if (n > capacity || !out) return PG_ERR_ARG;
    memcpy(out, h, n);
    return PG_OK;
}

int pg_format_rows(int kind, const pg_results *res, const pg_row_labels *labels, int n_threads, char *out,
                   size_t capacity, size_t *n_bytes) {
    return pg_format_rows_ex(kind, res, labels, 0, 0, n_threads, out, capacity, n_bytes);
}

int pg_format_rows_ex(int kind, const pg_results *res, const pg_row_labels *labels, int flags, int n_pools,
                      int n_threads, char *out, size_t capacity, size_t *n_bytes) {
    if (!res || !labels || !labels->positions) return PG_ERR_ARG;
    if (!((kind >= PG_KIND_OLS && kind <= PG_KIND_FISHER) || (kind >= PG_KIND_MLE && kind <= PG_KIND_GWALPHA_ML))) return PG_ERR_ARG;
    if (labels->text ? !labels->line_offsets : (!labels->chr_names || !labels->chr_index)) return PG_ERR_ARG;
    const int64_t L = res->n_loci;
    if (L > 0 && (!res->meta || !res->stats)) return PG_ERR_ARG;
    int T = n_threads < 1 ? 1 : n_threads;
    if ((int64_t)T > (L + 4095) / 4096) T = (int)((L + 4095) / 4096);  // not worth a thread below 4,096 loci
    if (T < 1) T = 1;
    std::vector<std::string> parts((size_t)T);
    if (T == 1) {
        format_range(kind, res, labels, 0, L, parts[0], flags, n_pools);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++) {
            const int64_t lo = L * t / T, hi = L * (t + 1) / T;
            th.emplace_back([=, &parts] { format_range(kind, res, labels, lo, hi, parts[(size_t)t], flags, n_pools); });
        }
        for (auto &x : th) x.join();
    }
    return emit(parts, out, capacity, n_bytes);
}

int pg_format_kinship_rows(int64_t n_columns, int k, const char *const *chromosome, const uint64_t *position,
                           const char *const *allele, const double *beta, const double *pval, int n_threads,
                           char *out, size_t capacity, size_t *n_bytes) {
    if (n_columns < 0 || k < 0 || (n_columns > 0 && (!chromosome || !position || !allele || !beta || !pval)))
        return PG_ERR_ARG;
    const int64_t rows = n_columns * k;
    int T = n_threads < 1 ? 1 : n_threads;
    if ((int64_t)T > (rows + 8191) / 8192) T = (int)((rows + 8191) / 8192);
    if (T < 1) T = 1;
    std::vector<std::string> parts((size_t)T);
    auto work = [&](int t) {
        std::string o;  // thread-local, handed over at the end (see format_range)
        const int64_t lo = rows * t / T, hi = rows * (t + 1) / T;
        o.reserve((size_t)(hi - lo) * 64 + 4096);
        char num[420];
        for (int64_t r = lo; r < hi; r++) {
            const int64_t j = r / n_columns, i = r - j * n_columns;  // phenotype outer, column inner
            o.append(chromosome[i]);
            o.push_back(',');
            o.append(num, (size_t)(put_u64(position[i], num) - num));
            o.push_back(',');
            o.append(allele[i]);
            o.append(",Pheno_");
            o.append(num, (size_t)(put_u64((uint64_t)j, num) - num));
            o.push_back(',');
            o.append(num, (size_t)(put_f64(beta[(size_t)j * n_columns + i], num) - num));
            o.push_back(',');
            o.append(num, (size_t)(put_f64(pval[(size_t)j * n_columns + i], num) - num));
            o.push_back('\n');
        }
        parts[(size_t)t] = std::move(o);
    };
    if (T == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++) th.emplace_back(work, t);
        for (auto &x : th) x.join();
    }
    return emit(parts, out, capacity, n_bytes);
}

// ---- sync2csv: the rows of SaveCsv::write_csv (src/base/sync.rs:1182-1262) -------------------------------------------
int pg_sort_loci(const pg_row_labels *labels, int64_t n_loci, int64_t *order_out) {
    if (!labels || !labels->positions || n_loci < 0 || (n_loci > 0 && !order_out)) return PG_ERR_ARG;
    if (labels->text ? !labels->line_offsets : (!labels->chr_names || !labels->chr_index)) return PG_ERR_ARG;
    const Labels L{labels};
    struct Key {
        const char *s;
        size_t n;
        uint64_t pos;
        int64_t l;
    };
    std::vector<Key> keys((size_t)n_loci);
    for (int64_t l = 0; l < n_loci; l++) {
        Key &k = keys[(size_t)l];
        L.chr(l, k.s, k.n);
        k.pos = labels->positions[l];
        k.l = l;
    }
    // `a.chromosome.cmp(&b.chromosome).then(a.position.cmp(&b.position))`, stable (src/base/sync.rs:1092-1101);
    // Rust compares strings byte-wise
    std::stable_sort(keys.begin(), keys.end(), [](const Key &a, const Key &b) {
        const int c = memcmp(a.s, b.s, a.n < b.n ? a.n : b.n);
        if (c != 0) return c < 0;
        if (a.n != b.n) return a.n < b.n;
        return a.pos < b.pos;
    });
    for (int64_t i = 0; i < n_loci; i++) order_out[i] = keys[(size_t)i].l;
    return PG_OK;
}

int pg_format_frequency_header(const char *const *pool_names, int n_pools, char *out, size_t capacity, size_t *n_bytes) {
    if (n_pools < 0 || (n_pools > 0 && !pool_names)) return PG_ERR_ARG;
    std::string h = "#chr,pos,allele,";  // src/base/sync.rs:1236-1240
    for (int i = 0; i < n_pools; i++) {
        if (i) h.push_back(',');
        h.append(pool_names[i]);
    }
    h.push_back('\n');
    if (n_bytes) *n_bytes = h.size();
    if (h.size() > capacity || !out) return PG_ERR_ARG;
    memcpy(out, h.data(), h.size());
    return PG_OK;
}

int pg_format_frequency_rows(int64_t n_columns, int n_pools, const double *columns, const int64_t *col_locus,
                             const uint8_t *col_allele, const pg_row_labels *labels, const int64_t *locus_order,
                             int64_t n_order, int n_threads, char *out, size_t capacity, size_t *n_bytes) {
    if (n_columns < 0 || n_pools < 1 || !labels || !labels->positions) return PG_ERR_ARG;
    if (n_columns > 0 && (!columns || !col_locus || !col_allele)) return PG_ERR_ARG;
    if (labels->text ? !labels->line_offsets : (!labels->chr_names || !labels->chr_index)) return PG_ERR_ARG;
    // the sequence of columns to print: as stored, or locus by locus in `locus_order`
    std::vector<int64_t> seq;
    if (locus_order) {
        int64_t max_l = -1;
        for (int64_t c = 0; c < n_columns; c++) {
            if (c && col_locus[c] < col_locus[c - 1]) return PG_ERR_ARG;  // columns must be grouped by ascending locus
            if (col_locus[c] > max_l) max_l = col_locus[c];
        }
        std::vector<int64_t> first((size_t)(max_l + 2), -1), count((size_t)(max_l + 2), 0);
        for (int64_t c = 0; c < n_columns; c++) {
            if (first[(size_t)col_locus[c]] < 0) first[(size_t)col_locus[c]] = c;
            count[(size_t)col_locus[c]]++;
        }
        seq.reserve((size_t)n_columns);
        for (int64_t i = 0; i < n_order; i++) {
            const int64_t l = locus_order[i];
            if (l < 0 || l > max_l || first[(size_t)l] < 0) continue;
            for (int64_t c = 0; c < count[(size_t)l]; c++) seq.push_back(first[(size_t)l] + c);
        }
    }
    const int64_t rows = locus_order ? (int64_t)seq.size() : n_columns;
    int T = n_threads < 1 ? 1 : n_threads;
    const int64_t per = std::max<int64_t>(1, 65536 / n_pools);  // rows worth a thread
    if ((int64_t)T > (rows + per - 1) / per) T = (int)((rows + per - 1) / per);
    if (T < 1) T = 1;
    std::vector<std::string> parts((size_t)T);
    const Labels L{labels};
    auto work = [&](int t) {
        std::string o;  // thread-local, handed over at the end (see format_range)
        const Rounder r6(6);
        const int64_t lo = rows * t / T, hi = rows * (t + 1) / T;
        o.reserve((size_t)(hi - lo) * (size_t)(24 + 9 * n_pools));
        char num[420];
        for (int64_t r = lo; r < hi; r++) {
            const int64_t c = locus_order ? seq[(size_t)r] : r;
            const int64_t l = col_locus[c];
            const char *cs;
            size_t cn;
            L.chr(l, cs, cn);
            o.append(cs, cn);
            o.push_back(',');
            o.append(num, (size_t)(put_u64(labels->positions[l], num) - num));
            o.push_back(',');
            o.push_back(kAlleleNames[col_allele[c] < 6 ? col_allele[c] : 4]);
            const double *col = columns + (size_t)c * n_pools;
            for (int i = 0; i < n_pools; i++) {
                o.push_back(',');
                o.append(num, (size_t)(r6.put(col[i], num) - num));  // parse_f64_roundup_and_own(x, 6), sync.rs:1247
            }
            o.push_back('\n');
        }
        parts[(size_t)t] = std::move(o);
    };
    if (T == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++) th.emplace_back(work, t);
        for (auto &x : th) x.join();
    }
    return emit(parts, out, capacity, n_bytes);
}

int pg_format_f64(double x, int n_digits, char *out, size_t capacity) {
    char tmp[420];
    char *e = (n_digits > 0) ? Rounder(n_digits).put(x, tmp) : put_f64(x, tmp);
    const size_t n = (size_t)(e - tmp);
    if (!out || n + 1 > capacity) return PG_ERR_ARG;
    memcpy(out, tmp, n);
    out[n] = 0;
    return (int)n;
}

}  // extern "C"
