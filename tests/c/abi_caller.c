/* A plain C caller of libmcevidence_b200.so: no Python, no torch -- what a C/C++/Fortran host program (or the patched
 * reference through any FFI) links against.  Reads a binary chain written by the test (int64 n, int64 cols, then
 * n*cols float64 in the CosmoMC column order: weight, -logL, parameters) and prints, one per line,
 *     single <logE...>      mce_evidence on one context
 *     group  <ndev> <nccl> <logE...>   mce_group_evidence over every GPU of the box
 *     sweep  <rows> <logE...>          mce_evidence_sweep, one line per batch size
 *     batch  <i> <logE...>            mce_group_evidence_batch over three interleaved sub-chains
 * Build: gcc -std=c99 -I include tests/c/abi_caller.c -L mcevidence_b200 -lmcevidence_b200 -lm -o abi_caller */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mcevidence_b200.h"

#define CHECK(call)                                                            \
    do {                                                                       \
        int rc_ = (call);                                                      \
        if (rc_ != MCE_OK) {                                                   \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, mce_last_error()); \
            return 1;                                                          \
        }                                                                      \
    } while (0)

static void print_vec(const char *tag, const double *v, int n) {
    printf("%s", tag);
    for (int i = 0; i < n; ++i) printf(" %.17g", v[i]);
    printf("\n");
}

int main(int argc, char **argv) {
    if (argc < 4) {
        fprintf(stderr, "usage: %s chain.bin ndim kmax\n", argv[0]);
        return 2;
    }
    const int ndim = atoi(argv[2]), kmax = atoi(argv[3]), nk = kmax - 1;
    FILE *fh = fopen(argv[1], "rb");
    if (!fh) return 2;
    int64_t n = 0, cols = 0;
    if (fread(&n, 8, 1, fh) != 1 || fread(&cols, 8, 1, fh) != 1) return 2;
    double *chain = (double *)malloc((size_t)n * cols * 8);
    if (fread(chain, 8, (size_t)n * cols, fh) != (size_t)n * cols) return 2;
    fclose(fh);
    double *logl = (double *)malloc((size_t)n * 8), *w = (double *)malloc((size_t)n * 8), sumw = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        w[i] = chain[i * cols];
        logl[i] = -chain[i * cols + 1];
        sumw += w[i];
    }
    if (mce_version() != MCE_ABI_VERSION) return 3;

    mce_problem p;
    memset(&p, 0, sizeof(p));
    p.query = chain + 2;          /* strided rows: the parameter block of every chain row, first ndim columns used */
    p.n_query = n;
    p.ld_query = cols;
    p.logl = logl;
    p.weights = w;
    p.d = ndim;
    p.kmax = kmax;
    p.log_sum_w = log(sumw);
    double logE[MCE_MAX_NEIGHBOURS], var[MCE_MAX_NEIGHBOURS];
    mce_result r;
    memset(&r, 0, sizeof(r));
    r.logE = logE;
    r.var_logE = var;

    mce_ctx *ctx = NULL;
    CHECK(mce_ctx_create(0, &ctx));
    CHECK(mce_evidence(ctx, &p, &r));
    print_vec("single", logE, nk);

    mce_group *g = NULL;
    CHECK(mce_group_create(0, NULL, &g));
    CHECK(mce_group_evidence(g, &p, &r));
    printf("group %d %d", mce_group_size(g), mce_group_uses_nccl(g));
    print_vec("", logE, nk);

    const int64_t rows[3] = {n / 5, n / 2, n};
    double sw_logE[3][MCE_MAX_NEIGHBOURS];
    mce_result sw[3];
    memset(sw, 0, sizeof(sw));
    for (int b = 0; b < 3; ++b) sw[b].logE = sw_logE[b];
    CHECK(mce_evidence_sweep(ctx, &p, 3, rows, sw));
    for (int b = 0; b < 3; ++b) {
        printf("sweep %lld", (long long)rows[b]);
        print_vec("", sw_logE[b], nk);
    }

    /* three independent sub-chains (every third row) as one batch over the group */
    mce_problem ps[3];
    mce_result rs[3];
    double b_logE[3][MCE_MAX_NEIGHBOURS];
    memset(rs, 0, sizeof(rs));
    double *bl[3], *bw[3];
    for (int i = 0; i < 3; ++i) {
        const int64_t m = (n - i + 2) / 3;
        bl[i] = (double *)malloc((size_t)m * 8);
        bw[i] = (double *)malloc((size_t)m * 8);
        double s = 0.0;
        for (int64_t j = 0; j < m; ++j) {
            bl[i][j] = logl[i + 3 * j];
            bw[i][j] = w[i + 3 * j];
            s += bw[i][j];
        }
        ps[i] = p;
        ps[i].query = chain + (size_t)i * cols + 2;
        ps[i].ld_query = 3 * cols;
        ps[i].n_query = m;
        ps[i].logl = bl[i];
        ps[i].weights = bw[i];
        ps[i].log_sum_w = log(s);
        rs[i].logE = b_logE[i];
    }
    CHECK(mce_group_evidence_batch(g, 3, ps, rs));
    for (int i = 0; i < 3; ++i) {
        printf("batch %d", i);
        print_vec("", b_logE[i], nk);
    }
    mce_group_destroy(g);
    mce_ctx_destroy(ctx);
    return 0;
}
