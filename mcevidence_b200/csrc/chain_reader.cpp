// Host-side chain ingestion: multi-threaded parser for CosmoMC-style
// whitespace-separated text chains, the `np.loadtxt` call of the reference (MCEvidence.py:564).
// Decimal -> double conversion is std::from_chars (correctly rounded, Eisel-Lemire fast path) with glibc
// strtod as the fallback for spellings from_chars rejects (leading '+', hexadecimal): the same values
// as Python's float() / np.loadtxt; `#` starts a comment; blank lines are skipped.
// Not on the GPU hot path, but after it text parsing dominates "seconds per logE" for grids
// of small chains (planck_mcevidence.py reads ~4.3e6 rows for 175 evidence calls).
#include "../../include/mcevidence_b200.h"

#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <charconv>
#include <system_error>
#include <thread>
#include <vector>

namespace {

struct Span { const char* b; const char* e; };

inline const char* skip_ws(const char* p, const char* e) {
    while (p < e && (*p == ' ' || *p == '\t' || *p == '\r')) ++p;
    return p;
}

// end of the data part of the line starting at p (stops at '\n' or '#'), and the start of the next line
inline void line_bounds(const char* p, const char* e, const char** data_end, const char** next) {
    const char* q = p;
    const char* de = nullptr;
    while (q < e && *q != '\n') {
        if (*q == '#' && !de) de = q;
        ++q;
    }
    *data_end = de ? de : q;
    *next = q < e ? q + 1 : e;
}

inline bool is_data_line(const char* p, const char* de) { return skip_ws(p, de) < de; }

int count_tokens(const char* p, const char* de) {
    int n = 0;
    p = skip_ws(p, de);
    while (p < de) {
        ++n;
        while (p < de && !(*p == ' ' || *p == '\t' || *p == '\r')) ++p;
        p = skip_ws(p, de);
    }
    return n;
}

}  // namespace

extern "C" {

int mce_read_chain_text(const char* path, int nthreads, double** data, int64_t* nrows, int64_t* ncols) {
    if (!path || !data || !nrows || !ncols) return MCE_ERR_ARG;
    *data = nullptr;
    *nrows = *ncols = 0;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return MCE_ERR_ARG;
    struct stat sb;
    if (fstat(fd, &sb) != 0 || sb.st_size <= 0) {
        close(fd);
        return MCE_ERR_ARG;
    }
    const size_t size = (size_t)sb.st_size;
    char* buf = (char*)malloc(size + 1);
    if (!buf) {
        close(fd);
        return MCE_ERR_NOMEM;
    }
    if (nthreads < 1) {
        nthreads = (int)std::thread::hardware_concurrency();
        if (nthreads < 1) nthreads = 1;
        if (nthreads > 32) nthreads = 32;
    }
    if ((size_t)nthreads > size / 65536 + 1) nthreads = (int)(size / 65536 + 1);

    // read the file in parallel byte ranges (the copy out of the page cache and the first touch of the buffer
    // are more than half of the wall time of a single-threaded read)
    {
        std::atomic<int> short_read(0);
        std::vector<std::thread> readers;
        for (int t = 0; t < nthreads; ++t) {
            readers.emplace_back([&, t]() {
                size_t lo = size * (size_t)t / (size_t)nthreads, hi = size * (size_t)(t + 1) / (size_t)nthreads;
                while (lo < hi) {
                    const ssize_t r = pread(fd, buf + lo, hi - lo, (off_t)lo);
                    if (r <= 0) {
                        short_read.store(1);
                        return;
                    }
                    lo += (size_t)r;
                }
            });
        }
        for (auto& th : readers) th.join();
        close(fd);
        if (short_read.load()) {
            free(buf);
            return MCE_ERR_ARG;
        }
    }
    buf[size] = '\0';   // strtod never runs past the buffer
    const char* end = buf + size;

    // split at line boundaries
    std::vector<Span> spans(nthreads);
    const char* cur = buf;
    for (int t = 0; t < nthreads; ++t) {
        const char* target = (t == nthreads - 1) ? end : buf + (size * (size_t)(t + 1)) / (size_t)nthreads;
        if (target < cur) target = cur;
        const char* q = target;
        while (q < end && *q != '\n') ++q;
        if (q < end) ++q;
        spans[t] = Span{cur, q};
        cur = q;
    }

    // pass 1: data lines per span, column count from the first data line of the file
    std::vector<int64_t> counts(nthreads, 0);
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t) {
        pool.emplace_back([&, t]() {
            int64_t n = 0;
            const char* p = spans[t].b;
            while (p < spans[t].e) {
                const char *de, *nx;
                line_bounds(p, spans[t].e, &de, &nx);
                if (is_data_line(p, de)) ++n;
                p = nx;
            }
            counts[t] = n;
        });
    }
    for (auto& th : pool) th.join();
    pool.clear();
    int64_t total = 0;
    std::vector<int64_t> first(nthreads, 0);
    for (int t = 0; t < nthreads; ++t) {
        first[t] = total;
        total += counts[t];
    }
    int cols = 0;
    for (const char* p = buf; p < end;) {
        const char *de, *nx;
        line_bounds(p, end, &de, &nx);
        if (is_data_line(p, de)) {
            cols = count_tokens(p, de);
            break;
        }
        p = nx;
    }
    if (total == 0 || cols == 0) {
        free(buf);
        return MCE_ERR_ARG;
    }
    double* out = (double*)malloc((size_t)total * (size_t)cols * sizeof(double));
    if (!out) {
        free(buf);
        return MCE_ERR_NOMEM;
    }

    // pass 2: parse
    std::atomic<int> bad(0);
    for (int t = 0; t < nthreads; ++t) {
        pool.emplace_back([&, t]() {
            int64_t row = first[t];
            const char* p = spans[t].b;
            while (p < spans[t].e && !bad.load(std::memory_order_relaxed)) {
                const char *de, *nx;
                line_bounds(p, spans[t].e, &de, &nx);
                if (is_data_line(p, de)) {
                    double* dst = out + (size_t)row * (size_t)cols;
                    const char* q = skip_ws(p, de);
                    int c = 0;
                    while (q < de) {
                        double v = 0.0;
                        const char* stop = q;
                        const std::from_chars_result fr = std::from_chars(q, de, v);
                        if (fr.ec == std::errc() && fr.ptr > q) {
                            stop = fr.ptr;
                        } else {   // '+1.0', hex floats, out-of-range spellings: glibc decides
                            char* sp = nullptr;
                            v = strtod(q, &sp);
                            stop = sp;
                        }
                        if (stop == q || stop > de || c >= cols ||
                            (stop < de && !(*stop == ' ' || *stop == '\t' || *stop == '\r'))) {   // not a number / ragged row
                            bad.store(1);
                            break;
                        }
                        dst[c++] = v;
                        q = skip_ws(stop, de);
                    }
                    if (c != cols) bad.store(1);
                    ++row;
                }
                p = nx;
            }
        });
    }
    for (auto& th : pool) th.join();
    free(buf);
    if (bad.load()) {
        free(out);
        return MCE_ERR_ARG;
    }
    *data = out;
    *nrows = total;
    *ncols = cols;
    return MCE_OK;
}

void mce_free_host(void* p) { free(p); }

}  // extern "C"
