// C ABI of libmcevidence_b200.so (declared in include/mcevidence_b200.h): context, workspace,
// host<->device staging and the stream-ordered pipeline K0 -> K1 -> K2 -> K3 -> (exact) -> K4.
// No CPU fallback: every compute entry point needs an sm_100 device.
#include "../../include/mcevidence_b200.h"
#include "common.cuh"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <thread>
#include <new>
#include <string>
#include <vector>

using namespace mce;

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            const int code_ = (e_ == cudaErrorMemoryAllocation) ? MCE_ERR_NOMEM : MCE_ERR_CUDA;        \
            return fail(code_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
        }                                                                                              \
    } while (0)

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return MCE_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        const size_t want = bytes + (bytes >> 3) + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            e = cudaMalloc(&p, bytes);
            if (e != cudaSuccess) {
                cudaGetLastError();
                return fail(MCE_ERR_NOMEM, "cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
            }
            cap = bytes;
            return MCE_OK;
        }
        cap = want;
        return MCE_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct DevScalars {
    // head: what the host reads back after a call (one copy of DEV_HEAD_BYTES)
    double out[3 * (MCE_MAX_NEIGHBOURS + 1)];
    double logj;
    int status;
    int nflagged;
    unsigned int rmax_bits;
    int pad;
    // device-only tail
    double mean[MCE_MAX_DIM];
    double W[MCE_MAX_DIM * MCE_MAX_DIM];
    double W_ref[MCE_MAX_DIM * MCE_MAX_DIM];
};
constexpr size_t DEV_HEAD_BYTES = offsetof(DevScalars, mean);

// pinned host buffer, grow-only
struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return MCE_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        if (cudaMallocHost(&p, bytes + 256) != cudaSuccess) {
            cudaGetLastError();
            p = nullptr;
            return fail(MCE_ERR_NOMEM, "cudaMallocHost of %zu bytes failed", bytes);
        }
        cap = bytes + 256;
        return MCE_OK;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

enum { EV_START, EV_H2D, EV_COV, EV_WHITEN, EV_FILTER, EV_RERANK, EV_EXACT, EV_REDUCE, EV_END, EV_COUNT };

}  // namespace

struct mce_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::mutex mu;
    int opt_tq = -1, opt_tr = -1, opt_packed = -1, opt_chunks = -1, opt_margin = -1, opt_sync = -1, opt_tail = -1;
    // engine of the filter: 0 = FFMA filter, 1 = tcgen05, 2 = chosen per problem (default); check = also evaluate every
    // pair exactly and record the worst |coarse - exact| / e; slack = percent of the default error allowance
    int opt_engine = 2, opt_tc_check = 0, opt_tc_swap = 0, opt_tc_slack = 100, opt_tc_stages = -1, opt_tc_perf = 0, opt_tc_timeline = 0, opt_tc_mma_warps = 2;
    double dbg_tc_ratio = 0.0, dbg_tc_pairs = 0.0;
    std::vector<float> dbg_tc_dump;   // tc_check: coarse scores of unit 0, tile 0 ([256 queries][128 references])
    DevBuf xq, xr, logl, w, yr, yq, tiles, tc_tiles, tc_debug, tc_dump, cand_idx, cand_th, dknn, idx, flagged, covpart, redpart, scal,
        sink, exact;
    DevBuf rows_q, rows_r;          // row maps of the query / reference set (burn-in, thinning, split on the device)
    DevBuf sweep_meta, sweep_out;   // batch-size sweep: boundaries / per-batch (max, sum, sum2) triples
    DevScalars* hscal = nullptr;    // pinned
    PinBuf hres;                    // pinned result heads of a batch (one per problem) / sweep triples
    PinBuf hmeta;                   // pinned staging of the sweep metadata
    cudaEvent_t ev[EV_COUNT] = {};
    // mce_evidence_batch: independent problems run concurrently on `lanes` child contexts (own stream + workspace each)
    std::vector<mce_ctx*> lanes;
    std::vector<cudaEvent_t> ev_pool;   // start / end events of the problems of a batch
    int opt_lanes = 8;
    int opt_tc_min_mpairs = 0;      // automatic engine choice: tcgen05 from this many million pairs on (0 = measured default)
    int opt_stage = 1;              // batches: pack the host arrays into a pinned arena on packer threads
    PinBuf hstage;
    int opt_sweep_seed = 1;         // sweep: one filter launch per chunk, seeded with the previous chunk's thresholds
    int share = 1;                  // the work-split model assumes sm_count / share SMs (several problems in flight)
};

namespace {

int validate(const mce_problem* p, bool need_evidence) {
    if (!p) return fail(MCE_ERR_ARG, "problem is NULL");
    if (!p->query) return fail(MCE_ERR_ARG, "problem.query is NULL");
    if (p->d < 1 || p->d > MCE_MAX_DIM) return fail(MCE_ERR_ARG, "ndim=%d outside [1, %d]", p->d, MCE_MAX_DIM);
    if (p->n_query < 1) return fail(MCE_ERR_ARG, "n_query=%lld must be positive", (long long)p->n_query);
    if (p->ld_query < p->d) return fail(MCE_ERR_ARG, "ld_query=%lld < ndim=%d", (long long)p->ld_query, p->d);
    if (p->ref) {
        if (p->n_ref < 1) return fail(MCE_ERR_ARG, "n_ref=%lld must be positive", (long long)p->n_ref);
        if (p->ld_ref < p->d) return fail(MCE_ERR_ARG, "ld_ref=%lld < ndim=%d", (long long)p->ld_ref, p->d);
    }
    const long long nref = p->ref ? p->n_ref : p->n_query;
    if (nref >= 0x7fffffffLL) return fail(MCE_ERR_ARG, "reference set of %lld rows exceeds 2^31-1", nref);
    if (p->q_begin < 0 || p->q_end < 0 || p->q_end > p->n_query || (p->q_end != 0 && p->q_begin >= p->q_end))
        return fail(MCE_ERR_ARG, "bad query row range [%lld, %lld) for n_query=%lld", (long long)p->q_begin,
                    (long long)p->q_end, (long long)p->n_query);
    if (need_evidence) {
        if (!p->logl || !p->weights) return fail(MCE_ERR_ARG, "problem.logl / problem.weights is NULL");
    }
    if (p->W_ref && !p->W) return fail(MCE_ERR_ARG, "W_ref given without W");
    if (p->query_rows || p->ref_rows) {
        if (p->flags & MCE_FLAG_DEVICE_PTRS) return fail(MCE_ERR_ARG, "row maps (query_rows / ref_rows) need host arrays");
        if (p->query_rows && p->n_query_base < 1) return fail(MCE_ERR_ARG, "query_rows given with n_query_base=%lld", (long long)p->n_query_base);
        if (p->ref_rows && (!p->ref || p->n_ref_base < 1))
            return fail(MCE_ERR_ARG, "ref_rows given without ref or with n_ref_base=%lld", (long long)p->n_ref_base);
    }
    return MCE_OK;
}

struct Sweep {               // batch-size sweep: one pass serves every batch size (MCEvidence.py:1039-1045)
    int nb = 0;
    const int64_t* rows = nullptr;   // S_b: batch b uses the first S_b rows of the query set
    mce_result* results = nullptr;   // nb results
};

struct Run {
    // inputs
    const mce_problem* p;
    bool evidence;
    int n_neighbors;  // knn mode only
    double* dknn_out;
    int64_t* idx_out;
    mce_result* res;
    mce_stats* stats;
    const Sweep* sweep = nullptr;
    DevScalars* hres = nullptr;      // pinned destination of the result head (default: the context's own)
    bool batch_lane = false;         // part of mce_evidence_batch: only the total time is measured
};

// State of an enqueued problem that is needed again after the stream has been synchronised.
struct Pending {
    Run r;
    bool use_tc = false, exact_only = false, cross = false;
    int d = 0, nk = 0, launches = 0;
    long long nq_all = 0, nq = 0, n_ref = 0;
    FilterConfig cfg{0, 0, 0, 0};
    int chunks = 1, list_len = 0, nfull = 0, nqblocks = 0, smem = 0, ctas = 0;
    DevScalars* hres = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // start / end events of a batch lane problem (else the context's own)
    std::vector<long long> sweep_rows;   // S_b per batch (sweep only)
    double* h_sweep_out = nullptr;       // pinned [nb][nk][3]
};

// Work decomposition model (calibrated on B200, profiles/r01_filter_variants.md).  Query blocks [0, nfull)
// scan the whole reference set as one unit (all full waves of blocks); blocks [nfull, nqblocks) are split
// into `chunks` reference ranges so that the last wave is short; small problems (less than one wave) are
// split throughout.  Every split restarts the top-L lists, and a list insertion costs as much as kappa
// pair evaluations, so
//     unit(m) = m + kappa L ln(m / L)                             pair-equivalents per query of an m-reference unit
//     T(c)    = nfull/slots unit(n_ref) + ceil(rem c / SMs) / ctas unit(n_ref / c)
// (the tail is counted per SM: an SM that receives one unit more than its neighbours finishes later)
// Returns T * (queries resident per SM) / (relative pair rate): comparable between tile shapes.
struct SplitPlan { int nfull, chunks; double cost; };
SplitPlan plan_split(long long nqblocks, int sm_count, int qpc, int ctas, long long n_ref, int list_len, int d,
                     bool wide, int max_chunks, double kappa_override = 0.0) {
    const long long slots = (long long)sm_count * ctas;
    const double kappa = kappa_override > 0.0 ? kappa_override : (wide ? 500.0 : 250.0) * 11.5 / (d + 1.5);
    auto unit = [&](double m) { return m + kappa * list_len * log(fmax(m / list_len, 2.0)); };
    const long long nf = (nqblocks / slots) * slots;
    const long long rem = nqblocks - nf;
    SplitPlan best{(int)nf, 1, 1e300};
    const double head = (double)(nf / slots) * unit((double)n_ref);
    // less than one wave of blocks: split at least until every CTA slot has a unit (a lone CTA on an SM runs
    // far below a quarter of the SM's rate, which the per-SM count above does not capture)
    int c_min = 1;
    if (nf == 0 && rem < slots) c_min = (int)((slots + rem - 1) / rem);
    if (c_min > max_chunks) c_min = max_chunks;
    for (int cc = c_min; cc <= max_chunks; ++cc) {
        const double tail = rem ? (double)((rem * cc + sm_count - 1) / sm_count) / ctas * unit((double)n_ref / cc) : 0.0;
        const double t = head + tail;
        if (t < best.cost) {
            best.cost = t;
            best.chunks = cc;
        }
        if (!rem) break;
    }
    best.cost *= (double)qpc * ctas / (wide ? 1.05 : 1.0);
    return best;
}

// Everything of one problem up to the device->host copy of its result head, enqueued on the context's stream without
// a host synchronisation (pageable host->device copies still wait for the stream, see the CUDA runtime).
int enqueue_problem(mce_ctx* c, const Run& r, Pending& pd) {
    const mce_problem& p = *r.p;
    int rc = validate(&p, r.evidence);
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    pd.r = r;
    const Sweep* sw = r.sweep;
    const bool devptr = (p.flags & MCE_FLAG_DEVICE_PTRS) != 0;
    const int f32 = (p.flags & MCE_FLAG_F32_SAMPLES) ? 1 : 0;   // query / ref are float32 arrays
    const size_t esz = f32 ? sizeof(float) : sizeof(double);
    const bool cross = p.ref != nullptr;
    const int d = p.d;
    const long long nq_all = p.n_query;
    long long smax = 0;
    if (sw) {
        // batch b = the first S_b rows of the query set; the largest batch bounds the search
        if (!r.evidence || sw->nb < 1 || sw->nb > 4096 || !sw->rows || !sw->results)
            return fail(MCE_ERR_ARG, "bad batch-size sweep arguments (1..4096 batches)");
        if (p.q_begin != 0 || p.q_end != 0) return fail(MCE_ERR_ARG, "a batch-size sweep takes no query row range");
        if (p.flags & (MCE_FLAG_DEVICE_PTRS | MCE_FLAG_EXACT_ONLY))
            return fail(MCE_ERR_ARG, "a batch-size sweep takes host arrays and the filtered search");
        for (int b = 0; b < sw->nb; ++b) {
            if (sw->rows[b] < 1 || sw->rows[b] > nq_all)
                return fail(MCE_ERR_ARG, "batch %d: %lld rows outside [1, n_query=%lld]", b, (long long)sw->rows[b], nq_all);
            if (sw->results[b].dknn || sw->results[b].idx)
                return fail(MCE_ERR_ARG, "batch %d: neighbour arrays are not returned by a sweep", b);
            if (!sw->results[b].logE) return fail(MCE_ERR_ARG, "results[%d].logE is NULL", b);
            if (sw->rows[b] > smax) smax = sw->rows[b];
        }
    }
    const bool sweep_auto = sw && !cross;          // prefixes of one set: chunked filter pass + per-batch re-rank
    const long long n_ref = cross ? p.n_ref : (sw ? smax : nq_all);
    const long long qb = sw ? 0 : p.q_begin, qe = sw ? smax : ((p.q_end == 0) ? nq_all : p.q_end);
    const long long nq = qe - qb;
    const int kmax = p.kmax < 2 ? 2 : p.kmax;

    // number of true neighbours and output shape
    int K, out_cols, out_col0, write_self;
    double* user_dknn;
    int64_t* user_idx;
    if (r.evidence) {
        user_dknn = r.res->dknn;
        user_idx = r.res->idx;
        const bool shaped = user_dknn || user_idx;
        if (shaped) {  // kneighbors-shaped (nq, kmax+1) output requested (MCEvidence.py:1104)
            out_cols = kmax + 1;
            out_col0 = cross ? 0 : 1;
            write_self = cross ? 0 : 1;
            K = cross ? kmax + 1 : kmax;
        } else {       // only what the estimator reads: neighbours 1..kmax-1 (auto) / 2..kmax (cross), F7
            K = cross ? kmax : kmax - 1;
            out_cols = K;
            out_col0 = 0;
            write_self = 0;
        }
    } else {
        user_dknn = r.dknn_out;
        user_idx = r.idx_out;
        out_cols = r.n_neighbors;
        out_col0 = cross ? 0 : 1;
        write_self = cross ? 0 : 1;
        K = cross ? r.n_neighbors : r.n_neighbors - 1;
        if (!user_dknn) return fail(MCE_ERR_ARG, "dknn output is NULL");
    }
    if (K < 1 || K > MCE_MAX_NEIGHBOURS)
        return fail(MCE_ERR_ARG, "number of neighbours %d outside [1, %d]", K, MCE_MAX_NEIGHBOURS);
    if (n_ref - (cross ? 0 : 1) < K)
        return fail(MCE_ERR_ARG, "Expected n_neighbors <= n_samples_fit: %d neighbours from %lld reference rows", K,
                    n_ref - (cross ? 0 : 1));
    if (sweep_auto)
        for (int b = 0; b < sw->nb; ++b)
            if (sw->rows[b] - 1 < K)
                return fail(MCE_ERR_ARG, "Expected n_neighbors <= n_samples_fit: %d neighbours from the %lld rows of batch %d",
                            K, (long long)sw->rows[b] - 1, b);

    int launches = 0;
    CU(cudaEventRecord(pd.ev0 ? pd.ev0 : c->ev[EV_START], st));

    // ---- inputs on the device ------------------------------------------------------------
    const void *dq = nullptr, *dr = nullptr;
    const long long *dq_rows = nullptr, *dr_rows = nullptr;
    const double *dlogl = nullptr, *dw = nullptr;
    long long ldq = p.ld_query, ldr = p.ld_ref;
    if (devptr) {
        dq = p.query;
        dr = p.ref;
        dlogl = p.logl;
        dw = p.weights;
    } else {
        // Strided rows (first d of ld columns, MCEvidence.py:902): a pitched copy from pageable memory
        // costs ~0.2 us per row, so whole rows are copied in one flat transfer and the kernels read
        // with the row stride; only very wide rows (ld > 16 d) are packed by a 2-D copy.
        auto upload = [&](DevBuf& buf, const void* src, long long n, long long ld_src, const void** dptr,
                          long long* ld_dev) -> int {
            if (ld_src == d || ld_src <= 16LL * d) {
                const size_t elems = (size_t)(n - 1) * ld_src + d;
                int rc2 = buf.ensure(elems * esz);
                if (rc2) return rc2;
                CU(cudaMemcpyAsync(buf.p, src, elems * esz, cudaMemcpyHostToDevice, st));
                *ld_dev = ld_src;
            } else {
                int rc2 = buf.ensure((size_t)n * d * esz);
                if (rc2) return rc2;
                CU(cudaMemcpy2DAsync(buf.p, (size_t)d * esz, src, (size_t)ld_src * esz, (size_t)d * esz, (size_t)n,
                                     cudaMemcpyHostToDevice, st));
                *ld_dev = d;
            }
            *dptr = buf.p;
            return MCE_OK;
        };
        // with a row map the array is the base the samples are gathered from (burn-in / thinning / split on the
        // device): the whole base goes up once, the covariance and whitening kernels follow the map as they load
        const long long nq_up = p.query_rows ? p.n_query_base : ((sw && p.W) ? smax : nq_all);
        rc = upload(c->xq, p.query, nq_up, p.ld_query, &dq, &ldq);
        if (rc) return rc;
        if (cross) {
            const bool same_base = p.query_rows && p.ref_rows && p.ref == p.query && p.ld_ref == p.ld_query &&
                                   p.n_ref_base == p.n_query_base;
            if (same_base) {      // the two halves of one split chain
                dr = dq;
                ldr = ldq;
            } else {
                rc = upload(c->xr, p.ref, p.ref_rows ? p.n_ref_base : n_ref, p.ld_ref, &dr, &ldr);
                if (rc) return rc;
            }
        }
        auto upload_rows = [&](DevBuf& buf, const int64_t* rows, long long n, long long n_base, const char* what,
                               const long long** dptr) -> int {
            for (long long i = 0; i < n; ++i)
                if (rows[i] < 0 || rows[i] >= n_base)
                    return fail(MCE_ERR_ARG, "%s[%lld]=%lld outside the %lld base rows", what, i, (long long)rows[i], n_base);
            int rc2 = buf.ensure((size_t)n * 8);
            if (rc2) return rc2;
            CU(cudaMemcpyAsync(buf.p, rows, (size_t)n * 8, cudaMemcpyHostToDevice, st));
            *dptr = buf.as<long long>();
            return MCE_OK;
        };
        if (p.query_rows) {
            rc = upload_rows(c->rows_q, p.query_rows, nq_all, p.n_query_base, "query_rows", &dq_rows);
            if (rc) return rc;
        }
        if (cross && p.ref_rows) {
            rc = upload_rows(c->rows_r, p.ref_rows, n_ref, p.n_ref_base, "ref_rows", &dr_rows);
            if (rc) return rc;
        }
        if (r.evidence) {
            rc = c->logl.ensure((size_t)nq * 8);
            if (rc) return rc;
            rc = c->w.ensure((size_t)nq * 8);
            if (rc) return rc;
            CU(cudaMemcpyAsync(c->logl.p, p.logl + qb, (size_t)nq * 8, cudaMemcpyHostToDevice, st));
            CU(cudaMemcpyAsync(c->w.p, p.weights + qb, (size_t)nq * 8, cudaMemcpyHostToDevice, st));
            dlogl = c->logl.as<double>() - qb;  // indexed by global row
            dw = c->w.as<double>() - qb;
        }
    }
    CU(cudaEventRecord(c->ev[EV_H2D], st));

    // ---- K0: whitening matrix -------------------------------------------------------------
    rc = c->scal.ensure(sizeof(DevScalars));
    if (rc) return rc;
    DevScalars* ds = c->scal.as<DevScalars>();
    CU(cudaMemsetAsync(&ds->status, 0, 4 * sizeof(int), st));  // status, nflagged, rmax_bits, pad
    const double *dW, *dWr, *dmu;
    if (!p.W) {
        const int grid = cov_grid_size(c->sm_count);
        const int npairs = d * (d + 1) / 2;
        rc = c->covpart.ensure((size_t)grid * (npairs > d ? npairs : d) * 8);
        if (rc) return rc;
        CovWorkspace ws{c->covpart.as<double>(), ds->mean, ds->W, &ds->logj, &ds->status};
        Segment a{dq, nq_all, ldq, f32, dq_rows};
        Segment b{cross ? dr : nullptr, cross ? n_ref : 0, ldr, f32, dr_rows};
        launch_covariance(a, b, d, ws, grid, st, &launches);
        dW = ds->W;
        dWr = ds->W;
        dmu = ds->mean;
    } else {
        CU(cudaMemcpyAsync(ds->W, p.W, (size_t)d * d * 8, cudaMemcpyHostToDevice, st));
        dW = ds->W;
        dWr = ds->W;
        if (p.W_ref) {
            CU(cudaMemcpyAsync(ds->W_ref, p.W_ref, (size_t)d * d * 8, cudaMemcpyHostToDevice, st));
            dWr = ds->W_ref;
        }
        if (p.mu) {
            CU(cudaMemcpyAsync(ds->mean, p.mu, (size_t)d * 8, cudaMemcpyHostToDevice, st));
            dmu = ds->mean;
        } else {
            dmu = nullptr;
        }
        CU(cudaMemcpyAsync(&ds->logj, &p.log_jacobian, 8, cudaMemcpyHostToDevice, st));
    }
    CU(cudaEventRecord(c->ev[EV_COV], st));

    // ---- K1: whitening on load ------------------------------------------------------------
    const int TJ = tile_refs(d);
    const int ntiles = (int)((n_ref + TJ - 1) / TJ);
    rc = c->yr.ensure((size_t)n_ref * d * 8);
    if (rc) return rc;
    rc = c->tiles.ensure((size_t)ntiles * (d + 1) * TJ * sizeof(float));
    if (rc) return rc;
    const double* yq_base;
    if (cross) {
        launch_whiten(Segment{dr, n_ref, ldr, f32, dr_rows}, d, dWr, dmu, c->yr.as<double>(), c->tiles.as<float>(), &ds->rmax_bits, st,
                      &launches);
        rc = c->yq.ensure((size_t)nq * d * 8);
        if (rc) return rc;
        launch_whiten(Segment{dq, nq_all, ldq, f32, dq_rows}.rows_from(qb, nq), d, dW, dmu, c->yq.as<double>(), nullptr, nullptr, st, &launches);
        yq_base = c->yq.as<double>() - qb * d;
    } else {
        launch_whiten(Segment{dq, n_ref, ldq, f32, dq_rows}, d, dW, dmu, c->yr.as<double>(), c->tiles.as<float>(), &ds->rmax_bits, st,
                      &launches);
        yq_base = c->yr.as<double>();
    }
    // engine 2 = automatic (the default): the tcgen05 engine where it was measured faster than the FFMA filter on B200
    // (profiles/r02_engine_sweep.json): ndim >= 4 and at least ~3e8 pairs (tools/small_probe.py, midsize_probe.py).  Below that the
    // 256-query x 128-reference tiles of one CTA per SM cannot fill the GPU, and for ndim <= 3 the TF32 allowance
    // admits too many false candidates.
    bool use_tc = false;
    if ((c->opt_engine == 1 || c->opt_engine == 2) && !(p.flags & MCE_FLAG_EXACT_ONLY)) {
        int q0 = 0, s0 = 0, c0 = 0, st0 = 0;
        const bool fits = tc_filter_query(d, K + (c->opt_margin >= 0 ? c->opt_margin : 2), c->opt_tc_stages, &q0, &s0, &c0, &st0) == 0;
        if (c->opt_engine == 1) {
            if (!fits) return fail(MCE_ERR_ARG, "no tensor-core filter for ndim=%d with %d neighbours", d, K);
            use_tc = true;
        } else {
            // an explicit FFMA tiling option (tests, tuning runs) keeps the FFMA filter
            const bool ffma_tuned = c->opt_tq > 0 || c->opt_tr > 0 || c->opt_packed >= 0 || c->opt_sync >= 0;
            const double pairs = (double)nq * (double)n_ref;
            const double min_pairs = c->opt_tc_min_mpairs > 0 ? 1e6 * c->opt_tc_min_mpairs : (d <= 6 ? 2e8 : 3e8);   // ndim <= 6: K = 8, one MMA per tile
            use_tc = fits && !ffma_tuned && d >= 4 && pairs >= min_pairs;
        }
    }
    if (use_tc) {
        rc = c->tc_tiles.ensure(tc_image_bytes(n_ref, d));
        if (rc) return rc;
        launch_tc_image(c->yr.as<double>(), c->tiles.as<float>(), n_ref, d, c->tc_tiles.as<float>(), st, &launches);
    }
    CU(cudaEventRecord(c->ev[EV_WHITEN], st));

    // ---- K2: FP32 filter ---------------------------------------------------------------------
    const bool exact_only = (p.flags & MCE_FLAG_EXACT_ONLY) != 0;
    const int margin = c->opt_margin >= 0 ? c->opt_margin : 2;
    int list_len = K + margin;
    FilterConfig cfg;
    filter_default_config(d, list_len, &cfg);
    if (c->opt_tq > 0) cfg.tq = c->opt_tq;
    if (c->opt_tr > 0) cfg.tr = c->opt_tr;
    if (c->opt_packed >= 0) cfg.packed = c->opt_packed;
    if (c->opt_sync >= 0) cfg.sync = c->opt_sync;
    int qpc = 0, smem = 0, ctas = 0, tc_stages = 0;
    // tiles of the engine in use: the split of the reference set is made at tile boundaries
    const int ftiles = use_tc ? (int)((n_ref + TC_TILE_REFS - 1) / TC_TILE_REFS) : ntiles;
    const int max_chunks = ftiles / 4 > 1 ? (ftiles / 4 > 64 ? 64 : ftiles / 4) : 1;
    if (use_tc) {
        if (tc_filter_query(d, list_len, c->opt_tc_stages, &qpc, &smem, &ctas, &tc_stages) != 0)
            return fail(MCE_ERR_ARG, "no tensor-core filter for ndim=%d list_len=%d", d, list_len);
        cfg = FilterConfig{1, TC_TILE_REFS, 0, 16};   // reported in the stats: mode 16 = tcgen05 engine
    } else if (c->opt_tq <= 0 && c->opt_tr <= 0 && c->opt_packed < 0 && c->opt_sync < 0 && d <= 16) {
        // Low ndim: four queries per thread (half the shared-memory traffic per FMA, two-row prefetch) is
        // 2-11 % faster when its 512-query blocks fill the GPU; the modelled cost decides (small inputs
        // fill the GPU with fewer reference splits using the 256-query blocks of the default).
        FilterConfig wide{4, 8, 1, 3};
        int qn = 0, sn = 0, cn = 0, q4 = 0, s4 = 0, c4 = 0;
        if (filter_query(d, cfg, list_len, &qn, &sn, &cn) == 0 && filter_query(d, wide, list_len, &q4, &s4, &c4) == 0) {
            const SplitPlan pn = plan_split((nq + qn - 1) / qn, c->sm_count, qn, cn, n_ref, list_len, d,
                                            false, max_chunks);
            const SplitPlan pw = plan_split((nq + q4 - 1) / q4, c->sm_count, q4, c4, n_ref, list_len, d,
                                            true, max_chunks);
            if (pw.cost < pn.cost) cfg = wide;
        } else {
            cudaGetLastError();
        }
    }
    if (!use_tc && filter_query(d, cfg, list_len, &qpc, &smem, &ctas) != 0) {
        cudaGetLastError();
        FilterConfig def;
        filter_default_config(d, list_len, &def);
        if (filter_query(d, def, list_len, &qpc, &smem, &ctas) != 0) {
            cudaGetLastError();
            return fail(MCE_ERR_ARG, "no filter kernel for ndim=%d tq=%d tr=%d packed=%d list_len=%d", d, cfg.tq, cfg.tr,
                        cfg.packed, list_len);
        }
        cfg = def;
    }
    const int nqblocks = (int)((nq + qpc - 1) / qpc);
    const long long nq_pad = (long long)nqblocks * qpc;
    // several problems in flight (mce_evidence_batch): plan for this problem's share of the SMs
    const int sm_eff = c->share > 1 ? (c->sm_count / c->share > 8 ? c->sm_count / c->share : 8) : c->sm_count;
    const long long slots = (long long)sm_eff * ctas;
    int chunks = 1, nfull = nqblocks;
    std::vector<int> chunk_tile0;      // sweep: chunk boundaries in tiles of the engine in use
    std::vector<long long> sw_off, sw_rows, sw_edge;
    std::vector<int> sw_nchunk;
    if (sweep_auto) {
        // Reference chunks end at the tile-aligned batch sizes: batch b merges the candidate lists of the chunks wholly
        // inside [0, S_b) and scans the partial tile [edge_b, S_b) directly in the re-rank.
        const int T = use_tc ? TC_TILE_REFS : TJ;
        std::vector<int> bnd;
        for (int b = 0; b < sw->nb; ++b)
            if (sw->rows[b] < smax && sw->rows[b] / T > 0) bnd.push_back((int)(sw->rows[b] / T));
        bnd.push_back(ftiles);
        std::sort(bnd.begin(), bnd.end());
        bnd.erase(std::unique(bnd.begin(), bnd.end()), bnd.end());
        chunk_tile0.push_back(0);
        for (int t : bnd) chunk_tile0.push_back(t);
        chunks = (int)bnd.size();
        nfull = 0;
        long long off = 0;
        for (int b = 0; b < sw->nb; ++b) {
            const long long sb = sw->rows[b];
            sw_off.push_back(off);
            sw_rows.push_back(sb);
            off += sb;
            if (sb == smax) {
                sw_nchunk.push_back(chunks);
                sw_edge.push_back(sb);
            } else {
                const int tb = (int)(sb / T);
                int nc = 0;
                while (nc < chunks && chunk_tile0[nc + 1] <= tb) ++nc;
                sw_nchunk.push_back(nc);
                sw_edge.push_back((long long)tb * T);
            }
        }
        sw_off.push_back(off);
        if (off >= 0x7fffffffLL) return fail(MCE_ERR_ARG, "batch-size sweep: %lld rows in total exceed 2^31-1", off);
    } else if (c->opt_chunks > 0) {
        chunks = c->opt_chunks;      // uniform split (tuning / tests)
        nfull = 0;
    } else if (c->opt_tail > 0) {
        chunks = c->opt_tail;        // whole-set units for the full waves, forced split of the remainder
        nfull = (int)((nqblocks / slots) * slots);
    } else {
        // tensor-core engine: a pair costs about 1/100 cycle per SM whatever ndim is, a slow-path visit a few hundred
        // warp cycles, i.e. relatively far more than in the FFMA filter (a guess until it is calibrated on hardware)
        const SplitPlan pl = plan_split(nqblocks, sm_eff, qpc, ctas, n_ref, list_len, d, cfg.tq >= 4, max_chunks,
                                        use_tc ? 1500.0 : 0.0);
        chunks = pl.chunks;
        nfull = pl.nfull;
    }
    int tiles_per_chunk = ftiles;
    if (!sweep_auto) {
        if (chunks > ftiles) chunks = ftiles;
        tiles_per_chunk = (ftiles + chunks - 1) / chunks;
        chunks = (ftiles + tiles_per_chunk - 1) / tiles_per_chunk;
        if (chunks == 1) nfull = nqblocks;
    }

    // candidate planes: chunk 0 for every query, chunks 1.. only for the queries of the split blocks
    const long long tail_q0 = (long long)nfull * qpc;
    const long long ntail_pad = (long long)(nqblocks - nfull) * qpc;
    const size_t cand_elems = (size_t)nq_pad + (size_t)(chunks - 1) * (size_t)ntail_pad;
    rc = c->cand_idx.ensure(cand_elems * list_len * sizeof(int));
    if (rc) return rc;
    rc = c->cand_th.ensure(cand_elems * sizeof(float));
    if (rc) return rc;
    const long long out_rows = sweep_auto ? sw_off.back() : nq;   // rows of the neighbour arrays
    rc = c->flagged.ensure((size_t)out_rows * sizeof(int));
    if (rc) return rc;
    // sweep metadata on the device: [off (nb+1) | rows (nb) | edge (nb)] long long, then [nchunk (nb) | chunk_tile0] int
    const long long *d_sw_off = nullptr, *d_sw_rows = nullptr, *d_sw_edge = nullptr;
    const int *d_sw_nchunk = nullptr, *d_chunk_tile0 = nullptr;
    if (sweep_auto) {
        const int nb = sw->nb;
        const size_t nll = (size_t)3 * nb + 1, nint = (size_t)nb + chunk_tile0.size();
        rc = c->sweep_meta.ensure(nll * 8 + nint * 4);
        if (rc) return rc;
        rc = c->hmeta.ensure(nll * 8 + nint * 4);
        if (rc) return rc;
        long long* hl = c->hmeta.as<long long>();
        memcpy(hl, sw_off.data(), (size_t)(nb + 1) * 8);
        memcpy(hl + nb + 1, sw_rows.data(), (size_t)nb * 8);
        memcpy(hl + 2 * nb + 1, sw_edge.data(), (size_t)nb * 8);
        int* hi = reinterpret_cast<int*>(hl + nll);
        memcpy(hi, sw_nchunk.data(), (size_t)nb * 4);
        memcpy(hi + nb, chunk_tile0.data(), chunk_tile0.size() * 4);
        CU(cudaMemcpyAsync(c->sweep_meta.p, c->hmeta.p, nll * 8 + nint * 4, cudaMemcpyHostToDevice, st));
        d_sw_off = c->sweep_meta.as<long long>();
        d_sw_rows = d_sw_off + nb + 1;
        d_sw_edge = d_sw_off + 2 * nb + 1;
        d_sw_nchunk = reinterpret_cast<const int*>(d_sw_off + nll);
        d_chunk_tile0 = d_sw_nchunk + nb;
    }

    if (!exact_only) {
        FilterParams fp;
        fp.yq = yq_base;
        fp.status = &ds->status;
        fp.tiles = c->tiles.as<float>();
        fp.cand_idx = c->cand_idx.as<int>();
        fp.cand_th = c->cand_th.as<float>();
        fp.q_begin = qb;
        fp.q_end = qe;
        fp.nq_pad = nq_pad;
        fp.ntail_pad = ntail_pad;
        fp.tail_q0 = tail_q0;
        fp.self_offset = 0;
        fp.self_exclude = cross ? 0 : 1;
        fp.ntiles = ftiles;
        fp.tiles_per_chunk = tiles_per_chunk;
        fp.ref_chunks = chunks;
        fp.nqblocks = nqblocks;
        fp.nfull = nfull;
        fp.list_len = list_len;
        fp.d = d;
        fp.chunk_tile0 = d_chunk_tile0;
        fp.chunk_base = 0;
        fp.seed_th = 0;
        // Sweep: one launch per reference chunk, in order, each covering every query block; chunk c starts from the final
        // thresholds of chunk c - 1 (a valid floor because the batches are nested prefixes), so the pass makes about as
        // many list insertions as one unchunked scan instead of restarting the lists at every batch size.
        const int nlaunch = (sweep_auto && c->opt_sweep_seed) ? chunks : 1;
        if (nlaunch > 1) {
            fp.ref_chunks = 1;
            fp.seed_th = 1;
        }
        if (use_tc) {
            // Bound on |coarse - exact| per pair, e = slack_qr |q| Rmax + slack_rr Rmax^2: both operands are
            // truncated to TF32 (2^-10 relative each, 2^-9 (1 + 2^-10) on a product, summed with Cauchy-Schwarz),
            // h = h_hi + h_lo loses at most 2^-20 |h|, and 2^-14 (|q| |r| + |h|) is allowed for the accumulation
            // inside the tensor core (FP32 would need K 2^-24); 1.25 on top.
            rc = c->tc_debug.ensure(2 * sizeof(unsigned int));
            if (rc) return rc;
            CU(cudaMemsetAsync(c->tc_debug.p, 0, 2 * sizeof(unsigned int), st));
            const double scale = 1.25 * (double)c->opt_tc_slack / 100.0;
            TcFilterParams tp;
            tp.f = fp;
            tp.f.tiles = c->tc_tiles.as<float>();
            tp.rmax_bits = &ds->rmax_bits;
            tp.slack_qr = (float)(scale * (ldexp(1.0, -9) * (1.0 + ldexp(1.0, -10)) + ldexp(1.0, -14)));
            tp.slack_rr = (float)(scale * 0.5 * (ldexp(1.0, -20) + ldexp(1.0, -14)));
            tp.debug = c->tc_debug.as<unsigned int>();
            tp.dump = nullptr;
            tp.timeline = nullptr;
            if (c->opt_tc_timeline && !c->opt_tc_check) {
                rc = c->tc_dump.ensure((size_t)TC_QUERIES * TC_TILE_REFS * sizeof(float));
                if (rc) return rc;
                CU(cudaMemsetAsync(c->tc_dump.p, 0, (size_t)TC_QUERIES * TC_TILE_REFS * sizeof(float), st));
                tp.timeline = c->tc_dump.as<long long>();   // 1024 tiles x 16 events x 8 bytes = the dump buffer
            }
            if (c->opt_tc_check) {
                rc = c->tc_dump.ensure((size_t)TC_QUERIES * TC_TILE_REFS * sizeof(float));
                if (rc) return rc;
                CU(cudaMemsetAsync(c->tc_dump.p, 0, (size_t)TC_QUERIES * TC_TILE_REFS * sizeof(float), st));
                tp.dump = c->tc_dump.as<float>();
            }
            tp.stages = tc_stages;
            tp.swap_offsets = c->opt_tc_swap;
            tp.check = c->opt_tc_check;
            tp.perf_mode = c->opt_tc_perf;
            tp.mma_warps = c->opt_tc_mma_warps;
            for (int cb = 0; cb < nlaunch; ++cb) {
                tp.f.chunk_base = cb;
                if (launch_tc_filter(tp, st, &launches) != 0)
                    return fail(MCE_ERR_CUDA, "tensor-core filter launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            }
        } else {
            for (int cb = 0; cb < nlaunch; ++cb) {
                fp.chunk_base = cb;
                if (launch_filter(fp, cfg, st, &launches) != 0)
                    return fail(MCE_ERR_CUDA, "filter launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            }
        }
    }
    CU(cudaEventRecord(c->ev[EV_FILTER], st));

    // ---- K3: FP64 re-rank + verification; exact re-scan of unverified rows --------------
    double* d_dknn;
    long long* d_idx = nullptr;
    if (devptr && user_dknn) {
        d_dknn = user_dknn;
    } else {
        rc = c->dknn.ensure((size_t)out_rows * out_cols * 8);
        if (rc) return rc;
        d_dknn = c->dknn.as<double>();
    }
    if (user_idx) {
        if (devptr) {
            d_idx = reinterpret_cast<long long*>(user_idx);
        } else {
            rc = c->idx.ensure((size_t)out_rows * out_cols * 8);
            if (rc) return rc;
            d_idx = c->idx.as<long long>();
        }
    }
    RerankParams rp;
    rp.yq = yq_base;
    rp.status = &ds->status;
    rp.yr = c->yr.as<double>();
    rp.cand_idx = c->cand_idx.as<int>();
    rp.cand_th = c->cand_th.as<float>();
    rp.rmax_bits = &ds->rmax_bits;
    rp.dknn = d_dknn;
    rp.idx = d_idx;
    rp.flagged = c->flagged.as<int>();
    rp.nflagged = &ds->nflagged;
    rp.q_begin = qb;
    rp.q_end = qe;
    rp.nq_pad = nq_pad;
    rp.ntail_pad = ntail_pad;
    rp.tail_q0 = tail_q0;
    rp.n_ref = n_ref;
    rp.self_offset = 0;
    rp.self_exclude = cross ? 0 : 1;
    rp.ref_chunks = chunks;
    rp.nfull_queries = (int)((long long)nfull * qpc < nq ? (long long)nfull * qpc : nq);
    rp.list_len = list_len;
    rp.d = d;
    rp.k = K;
    rp.out_ld = out_cols;
    rp.out_col0 = out_col0;
    rp.write_self_col = write_self;
    rp.flag_all = exact_only ? 1 : 0;
    rp.sweep_off = d_sw_off;
    rp.sweep_rows = d_sw_rows;
    rp.sweep_edge = d_sw_edge;
    rp.sweep_nchunk = d_sw_nchunk;
    rp.sweep_nb = sweep_auto ? sw->nb : 0;
    rp.sweep_total = sweep_auto ? out_rows : 0;
    launch_rerank(rp, st, &launches);
    CU(cudaEventRecord(c->ev[EV_RERANK], st));
    const int exact_grid = c->sm_count * 4;
    rc = c->exact.ensure(exact_scratch_bytes(exact_grid));
    if (rc) return rc;
    launch_exact_rescan(rp, exact_grid, c->exact.p, st, &launches);
    CU(cudaEventRecord(c->ev[EV_EXACT], st));

    // ---- K4: epilogue -------------------------------------------------------------------------
    const int nk = kmax - 1;
    if (r.evidence) {
        ReduceParams ep;
        ep.logl = dlogl;
        ep.weights = dw;
        ep.ld = out_cols;
        ep.col0 = out_col0;
        ep.first_nb = cross ? 2 : 1;
        ep.nk = nk;
        ep.d = d;
        if (!sw) {
            const int rgrid = reduce_grid_size(nq);
            rc = c->redpart.ensure((size_t)rgrid * nk * 3 * 8);
            if (rc) return rc;
            ep.dknn = d_dknn;
            ep.block_partial = c->redpart.as<double>();
            ep.out = ds->out;
            ep.q_begin = qb;
            ep.q_end = qe;
            launch_reduce(ep, rgrid, st, &launches);
        } else {
            // one reduction per batch size over its rows: rows [off_b, off_b + S_b) of the sweep's neighbour array
            // (auto mode), or the first S_b rows of the one search against s2 (cross mode)
            rc = c->redpart.ensure((size_t)reduce_grid_size(smax) * nk * 3 * 8);
            if (rc) return rc;
            rc = c->sweep_out.ensure((size_t)sw->nb * nk * 3 * 8);
            if (rc) return rc;
            for (int b = 0; b < sw->nb; ++b) {
                const long long sb = sw->rows[b];
                ep.dknn = d_dknn + (sweep_auto ? sw_off[b] : 0) * out_cols;
                ep.block_partial = c->redpart.as<double>();
                ep.out = c->sweep_out.as<double>() + (size_t)b * nk * 3;
                ep.q_begin = 0;
                ep.q_end = sb;
                launch_reduce(ep, reduce_grid_size(sb), st, &launches);
            }
        }
    }
    CU(cudaEventRecord(c->ev[EV_REDUCE], st));

    // ---- results back ---------------------------------------------------------------------------
    DevScalars* hres = r.hres ? r.hres : c->hscal;
    CU(cudaMemcpyAsync(hres, ds, DEV_HEAD_BYTES, cudaMemcpyDeviceToHost, st));
    if (sw) {
        rc = c->hres.ensure((size_t)sw->nb * nk * 3 * 8);
        if (rc) return rc;
        pd.h_sweep_out = c->hres.as<double>();
        CU(cudaMemcpyAsync(pd.h_sweep_out, c->sweep_out.p, (size_t)sw->nb * nk * 3 * 8, cudaMemcpyDeviceToHost, st));
    }
    if (!devptr) {
        if (user_dknn) CU(cudaMemcpyAsync(user_dknn, d_dknn, (size_t)nq * out_cols * 8, cudaMemcpyDeviceToHost, st));
        if (user_idx) CU(cudaMemcpyAsync(user_idx, d_idx, (size_t)nq * out_cols * 8, cudaMemcpyDeviceToHost, st));
    }
    CU(cudaEventRecord(pd.ev1 ? pd.ev1 : c->ev[EV_END], st));

    pd.use_tc = use_tc;
    pd.exact_only = exact_only;
    pd.cross = cross;
    pd.d = d;
    pd.nk = nk;
    pd.launches = launches;
    pd.nq_all = nq_all;
    pd.nq = nq;
    pd.n_ref = n_ref;
    pd.cfg = cfg;
    pd.chunks = chunks;
    pd.list_len = list_len;
    pd.nfull = nfull;
    pd.nqblocks = nqblocks;
    pd.smem = smem;
    pd.ctas = ctas;
    pd.hres = hres;
    if (sw) pd.sweep_rows.assign(sw->rows, sw->rows + sw->nb);
    return MCE_OK;
}

// After the stream has been synchronised: diagnostics, error state, statistics and the host arithmetic of the result.
int complete_problem(mce_ctx* c, Pending& pd) {
    const Run& r = pd.r;
    const mce_problem& p = *r.p;
    cudaStream_t st = c->stream;
    const bool use_tc = pd.use_tc, exact_only = pd.exact_only;
    const int d = pd.d, nk = pd.nk;
    if (use_tc && c->opt_tc_timeline && !c->opt_tc_check && !exact_only) {
        c->dbg_tc_dump.resize((size_t)TC_QUERIES * TC_TILE_REFS);
        CU(cudaMemcpy(c->dbg_tc_dump.data(), c->tc_dump.p, c->dbg_tc_dump.size() * sizeof(float), cudaMemcpyDeviceToHost));
    }
    if (use_tc && c->opt_tc_check && !exact_only) {
        unsigned int dbg[2] = {0, 0};
        CU(cudaMemcpy(dbg, c->tc_debug.p, sizeof(dbg), cudaMemcpyDeviceToHost));
        float ratio;
        memcpy(&ratio, &dbg[0], 4);
        c->dbg_tc_ratio = ratio;
        c->dbg_tc_pairs = dbg[1];
        c->dbg_tc_dump.resize((size_t)TC_QUERIES * TC_TILE_REFS);
        CU(cudaMemcpy(c->dbg_tc_dump.data(), c->tc_dump.p, c->dbg_tc_dump.size() * sizeof(float), cudaMemcpyDeviceToHost));
    }
    if (use_tc && !exact_only) {
        unsigned int stuck[4] = {0u, 0u, 0u, 0u};
        if (tc_read_stuck(stuck, st) != 0) return fail(MCE_ERR_CUDA, "reading the tcgen05 filter's wait record failed");
        if (stuck[0] != 0u)
            return fail(MCE_ERR_INTERNAL, "tcgen05 filter: a %s wait made no progress for 0.5 s and gave up (a=0x%x b=0x%x block=%u); "
                        "results discarded", stuck[0] == 1u ? "barrier" : stuck[0] == 2u ? "queue-full" : "queue-slot", stuck[1], stuck[2], stuck[3]);
    }
    const DevScalars* h = pd.hres;
    if (h->status != 0)
        return fail(MCE_ERR_NOT_POSDEF,
                    "sample covariance of the %d selected parameters is not positive definite "
                    "(reference: negative eigenvalue warning, MCEvidence.py:867)", d);

    mce_stats sst;
    memset(&sst, 0, sizeof(sst));
    float ms = 0.f;
    auto el = [&](int a, int b) { cudaEventElapsedTime(&ms, c->ev[a], c->ev[b]); return (double)ms; };
    if (pd.ev0) {   // batch lane: the lane's stage events belong to its last problem; only the total is per problem
        cudaEventElapsedTime(&ms, pd.ev0, pd.ev1);
        sst.ms_total = ms;
    } else {
        sst.ms_total = el(EV_START, EV_END);
        sst.ms_h2d = el(EV_START, EV_H2D);
        sst.ms_cov = el(EV_H2D, EV_COV);
        sst.ms_whiten = el(EV_COV, EV_WHITEN);
        sst.ms_filter = el(EV_WHITEN, EV_FILTER);
        sst.ms_rerank = el(EV_FILTER, EV_RERANK);
        sst.ms_exact = el(EV_RERANK, EV_EXACT);
        sst.ms_reduce = el(EV_EXACT, EV_REDUCE);
        sst.ms_d2h = el(EV_REDUCE, EV_END);
    }
    sst.pairs = pd.nq * pd.n_ref;
    sst.rows_rescanned = h->nflagged;
    sst.launches = pd.launches;
    sst.filter_tq = pd.cfg.tq;
    sst.filter_tr = pd.cfg.tr;
    sst.filter_packed = pd.cfg.packed;
    sst.filter_mode = pd.cfg.sync;
    sst.ref_chunks = pd.chunks;
    sst.list_len = pd.list_len;
    sst.filter_grid = pd.nfull + (pd.nqblocks - pd.nfull) * pd.chunks;
    sst.filter_smem = pd.smem;
    sst.filter_ctas_per_sm = pd.ctas;
    memcpy(&sst.rmax, &h->rmax_bits, 4);

    if (r.sweep) {
        for (int b = 0; b < r.sweep->nb; ++b) {
            mce_result* res = &r.sweep->results[b];
            const double* out = pd.h_sweep_out + (size_t)b * nk * 3;
            res->stats = sst;
            res->log_jacobian = h->logj;
            if ((p.flags & MCE_FLAG_PARTIALS) && res->partials) memcpy(res->partials, out, (size_t)nk * 3 * 8);
            // S = the rows of this batch, SumW = all of s1 (MCEvidence.py:1123-1131, 1126)
            const int rc = mce_finalize(out, nk, d, pd.cross ? 1 : 0, pd.sweep_rows[b], p.log_sum_w, h->logj, p.log_prior_volume,
                                        res->logE, res->var_logE);
            if (rc) return rc;
        }
    } else if (r.evidence) {
        mce_result* res = r.res;
        res->stats = sst;
        res->log_jacobian = h->logj;
        if ((p.flags & MCE_FLAG_PARTIALS) && res->partials) memcpy(res->partials, h->out, (size_t)nk * 3 * 8);
        if (res->logE) {
            const int rc = mce_finalize(h->out, nk, d, pd.cross ? 1 : 0, pd.nq_all, p.log_sum_w, h->logj,
                                        p.log_prior_volume, res->logE, res->var_logE);
            if (rc) return rc;
        }
    } else if (r.stats) {
        *r.stats = sst;
    }
    return MCE_OK;
}

// one problem on the context's own stream: enqueue, wait, complete
int run_problem(mce_ctx* c, const Run& r) {
    Pending pd;
    int rc = enqueue_problem(c, r, pd);
    if (rc) {
        cudaStreamSynchronize(c->stream);   // nothing of a failed call stays in flight
        cudaGetLastError();
        return rc;
    }
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    return complete_problem(c, pd);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
extern "C" {

int mce_version(void) { return MCE_ABI_VERSION; }

const char* mce_last_error(void) { return g_err; }

int mce_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int mce_ctx_create(int device, mce_ctx** out) {
    if (!out) return fail(MCE_ERR_ARG, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(MCE_ERR_CUDA, "no CUDA device available (%s); mcevidence_b200 has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= n) return fail(MCE_ERR_ARG, "device %d outside [0, %d)", device, n);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(MCE_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major,
                    prop.minor);
    mce_ctx* c = new (std::nothrow) mce_ctx();
    if (!c) return fail(MCE_ERR_NOMEM, "out of host memory");
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (const char* eng = getenv("MCE_FILTER_ENGINE"))
        if (eng[0] >= '0' && eng[0] <= '2' && eng[1] == '\0') c->opt_engine = eng[0] - '0';
    if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMallocHost(reinterpret_cast<void**>(&c->hscal), sizeof(DevScalars)) != cudaSuccess) {
        const char* msg = cudaGetErrorString(cudaGetLastError());
        delete c;
        return fail(MCE_ERR_CUDA, "context setup failed: %s", msg);
    }
    c->stream = c->own_stream;
    for (int i = 0; i < EV_COUNT; ++i) {
        if (cudaEventCreate(&c->ev[i]) != cudaSuccess) {
            mce_ctx_destroy(c);
            return fail(MCE_ERR_CUDA, "cudaEventCreate failed");
        }
    }
    *out = c;
    return MCE_OK;
}

void mce_ctx_destroy(mce_ctx* c) {
    if (!c) return;
    for (mce_ctx* l : c->lanes) mce_ctx_destroy(l);
    c->lanes.clear();
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    c->ev_pool.clear();
    cudaSetDevice(c->device);
    if (c->own_stream) cudaStreamSynchronize(c->own_stream);
    DevBuf* bufs[] = {&c->xq, &c->xr, &c->logl, &c->w, &c->yr, &c->yq, &c->tiles, &c->tc_tiles, &c->tc_debug, &c->tc_dump, &c->cand_idx, &c->cand_th,
                      &c->dknn, &c->idx, &c->flagged, &c->covpart, &c->redpart, &c->scal, &c->sink, &c->exact, &c->sweep_meta,
                      &c->sweep_out, &c->rows_q, &c->rows_r};
    for (DevBuf* b : bufs) b->release();
    c->hres.release();
    c->hmeta.release();
    c->hstage.release();
    for (int i = 0; i < EV_COUNT; ++i)
        if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    if (c->hscal) cudaFreeHost(c->hscal);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

int mce_ctx_set_stream(mce_ctx* c, void* cuda_stream) {
    if (!c) return fail(MCE_ERR_ARG, "ctx is NULL");
    std::lock_guard<std::mutex> lk(c->mu);
    c->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : c->own_stream;
    return MCE_OK;
}

int mce_ctx_set_option(mce_ctx* c, const char* name, int value) {
    if (!c || !name) return fail(MCE_ERR_ARG, "ctx or name is NULL");
    std::lock_guard<std::mutex> lk(c->mu);
    if (!strcmp(name, "filter_tq")) c->opt_tq = value;
    else if (!strcmp(name, "filter_tr")) c->opt_tr = value;
    else if (!strcmp(name, "filter_packed")) c->opt_packed = value;
    else if (!strcmp(name, "ref_chunks")) c->opt_chunks = value;
    else if (!strcmp(name, "list_margin")) c->opt_margin = value;
    else if (!strcmp(name, "filter_sync")) {
        if (value == 0) return fail(MCE_ERR_ARG, "filter_sync 0 (free-running warps) is no longer built: use 1, 2 or 3, or -1 for automatic");
        c->opt_sync = value;
    }
    else if (!strcmp(name, "tail_chunks")) c->opt_tail = value;
    else if (!strcmp(name, "sweep_seed")) c->opt_sweep_seed = value != 0;
    else if (!strcmp(name, "batch_stage")) c->opt_stage = value != 0;
    else if (!strcmp(name, "tc_min_mpairs")) c->opt_tc_min_mpairs = value > 0 ? value : 0;
    else if (!strcmp(name, "batch_lanes")) {
        if (value < 1 || value > 32) return fail(MCE_ERR_ARG, "batch_lanes must be in [1, 32]");
        c->opt_lanes = value;
    }
    else if (!strcmp(name, "filter_engine")) {
        if (value < 0 || value > 2) return fail(MCE_ERR_ARG, "filter_engine must be 0 (FFMA), 1 (tcgen05) or 2 (automatic)");
        c->opt_engine = value;
    } else if (!strcmp(name, "tc_check")) c->opt_tc_check = value != 0;
    else if (!strcmp(name, "tc_swap_offsets")) c->opt_tc_swap = value & 3;   // bit 0: exchange the offsets, bit 1: version field 0
    else if (!strcmp(name, "tc_stages")) c->opt_tc_stages = value;
    else if (!strcmp(name, "tc_timeline")) c->opt_tc_timeline = value != 0;
    else if (!strcmp(name, "tc_mma_warps")) c->opt_tc_mma_warps = value == 1 ? 1 : 2;
    else if (!strcmp(name, "tc_perf_mode")) c->opt_tc_perf = (value >= 1 && value <= 3) ? value : 0;
    else if (!strcmp(name, "tc_slack_percent")) {
        if (value < 1) return fail(MCE_ERR_ARG, "tc_slack_percent must be positive");
        c->opt_tc_slack = value;
    }
    else return fail(MCE_ERR_ARG, "unknown option '%s'", name);
    return MCE_OK;
}

int mce_ctx_get_debug(mce_ctx* c, const char* name, double* value) {
    if (!c || !name || !value) return fail(MCE_ERR_ARG, "ctx, name or value is NULL");
    std::lock_guard<std::mutex> lk(c->mu);
    if (!strcmp(name, "tc_check_ratio")) *value = c->dbg_tc_ratio;
    else if (!strcmp(name, "tc_check_pairs")) *value = c->dbg_tc_pairs;
    else if (!strcmp(name, "filter_engine")) *value = c->opt_engine;
    else return fail(MCE_ERR_ARG, "unknown debug value '%s'", name);
    return MCE_OK;
}

int mce_ctx_get_debug_array(mce_ctx* c, const char* name, float* out, int64_t capacity, int64_t* n) {
    if (!c || !name || !n) return fail(MCE_ERR_ARG, "ctx, name or n is NULL");
    std::lock_guard<std::mutex> lk(c->mu);
    if (strcmp(name, "tc_dump")) return fail(MCE_ERR_ARG, "unknown debug array '%s'", name);
    *n = (int64_t)c->dbg_tc_dump.size();
    if (out) {
        if (capacity < *n) return fail(MCE_ERR_ARG, "capacity %lld < %lld", (long long)capacity, (long long)*n);
        memcpy(out, c->dbg_tc_dump.data(), (size_t)*n * sizeof(float));
    }
    return MCE_OK;
}

int mce_evidence(mce_ctx* c, const mce_problem* problem, mce_result* result) {
    if (!c) return fail(MCE_ERR_ARG, "ctx is NULL");
    if (!result) return fail(MCE_ERR_ARG, "result is NULL");
    if (!result->logE && !((problem ? problem->flags : 0) & MCE_FLAG_PARTIALS))
        return fail(MCE_ERR_ARG, "result.logE is NULL");
    std::lock_guard<std::mutex> lk(c->mu);
    Run r{problem, true, 0, nullptr, nullptr, result, nullptr};
    return run_problem(c, r);
}

// Independent problems (the (model, data, chain) loop of planck_mcevidence.py:306-359) overlap on up to "batch_lanes"
// child contexts: each lane has its own stream and workspace, problems are dealt largest first and enqueued without
// waiting for results; one synchronisation per lane at the end.  A problem is planned for its share of the SMs, so small
// chains are not cut into many short units just to fill the GPU on their own.
int mce_evidence_batch(mce_ctx* c, int n, const mce_problem* problems, mce_result* results) {
    if (!c) return fail(MCE_ERR_ARG, "ctx is NULL");
    if (n < 0 || (n > 0 && (!problems || !results))) return fail(MCE_ERR_ARG, "bad batch arguments");
    std::lock_guard<std::mutex> lk(c->mu);
    for (int i = 0; i < n; ++i)
        if (!results[i].logE) return fail(MCE_ERR_ARG, "results[%d].logE is NULL", i);
    if (n == 0) return MCE_OK;
    auto named = [&](int i, int rc) {
        char tmp[sizeof(g_err)];
        snprintf(tmp, sizeof(tmp), "%s", g_err);
        return fail(rc, "problem %d: %s", i, tmp);
    };
    const int nl = c->opt_lanes < n ? c->opt_lanes : n;
    if (nl <= 1) {
        for (int i = 0; i < n; ++i) {
            Run r{&problems[i], true, 0, nullptr, nullptr, &results[i], nullptr};
            const int rc = run_problem(c, r);
            if (rc) return named(i, rc);
        }
        return MCE_OK;
    }
    CU(cudaSetDevice(c->device));
    while ((int)c->lanes.size() < nl) {
        mce_ctx* l = nullptr;
        const int rc = mce_ctx_create(c->device, &l);
        if (rc) return rc;
        c->lanes.push_back(l);
    }
    while ((int)c->ev_pool.size() < 2 * n) {
        cudaEvent_t e;
        CU(cudaEventCreate(&e));
        c->ev_pool.push_back(e);
    }
    int rc = c->hres.ensure((size_t)n * sizeof(DevScalars));
    if (rc) return rc;
    // largest first, round-robin over the lanes (the enqueue order interleaves the lanes)
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    auto cost = [&](int i) {
        const mce_problem& q = problems[i];
        return (double)q.n_query * (double)(q.ref ? q.n_ref : q.n_query);
    };
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost(a) > cost(b); });
    for (int l = 0; l < nl; ++l) {
        mce_ctx* lane = c->lanes[l];
        lane->opt_tq = c->opt_tq; lane->opt_tr = c->opt_tr; lane->opt_packed = c->opt_packed; lane->opt_chunks = c->opt_chunks;
        lane->opt_margin = c->opt_margin; lane->opt_sync = c->opt_sync; lane->opt_tail = c->opt_tail;
        lane->opt_engine = c->opt_engine; lane->opt_tc_slack = c->opt_tc_slack; lane->opt_tc_stages = c->opt_tc_stages;
        lane->opt_tc_mma_warps = c->opt_tc_mma_warps;
        lane->opt_tc_min_mpairs = c->opt_tc_min_mpairs;
        lane->share = nl;
    }
    // Host staging.  A copy from pageable memory waits for the lane's stream and goes through the driver's own bounce
    // buffer, which made the host side of a batch as long as its device side.  A few packer threads therefore copy every
    // problem's arrays -- only the d columns that are used, dense -- into one pinned arena, in enqueue order and ahead
    // of the enqueuing thread; the uploads are then true asynchronous copies that overlap the kernels of other lanes.
    std::vector<mce_problem> staged(problems, problems + n);
    std::vector<size_t> arena_off(n, 0);
    size_t arena_bytes = 0;
    const size_t arena_cap = (size_t)2 << 30;
    bool stage = c->opt_stage != 0;
    for (int i = 0; i < n && stage; ++i) {
        const mce_problem& q = problems[i];
        if ((q.flags & MCE_FLAG_DEVICE_PTRS) || q.query_rows || q.ref_rows || validate(&q, true) != MCE_OK) stage = false;
        const size_t esz = (q.flags & MCE_FLAG_F32_SAMPLES) ? 4 : 8;
        arena_off[i] = arena_bytes;
        arena_bytes += (((size_t)q.n_query * q.d + (q.ref ? (size_t)q.n_ref * q.d : 0)) * esz + (size_t)q.n_query * 16 + 255) & ~(size_t)255;
        if (arena_bytes > arena_cap) stage = false;
    }
    std::vector<std::atomic<int>> ready(stage ? n : 0);
    std::vector<std::thread> packers;
    std::atomic<int> next_job{0};
    if (stage) {
        rc = c->hstage.ensure(arena_bytes);
        if (rc) return rc;
        unsigned char* arena = c->hstage.as<unsigned char>();
        for (int i = 0; i < n; ++i) ready[i].store(0, std::memory_order_relaxed);
        auto pack_rows = [](unsigned char* dst, const void* src, long long rows, long long ld, int d, size_t esz) {
            const unsigned char* s0 = static_cast<const unsigned char*>(src);
            if (ld == d) {
                memcpy(dst, s0, (size_t)rows * d * esz);
            } else {
                for (long long r = 0; r < rows; ++r) memcpy(dst + (size_t)r * d * esz, s0 + (size_t)r * ld * esz, (size_t)d * esz);
            }
        };
        for (int i = 0; i < n; ++i) {      // where every array of problem i will live
            mce_problem& q = staged[i];
            const size_t esz = (q.flags & MCE_FLAG_F32_SAMPLES) ? 4 : 8;
            unsigned char* base = arena + arena_off[i];
            unsigned char* p_ref = base + (size_t)q.n_query * q.d * esz;
            unsigned char* p_logl = p_ref + (q.ref ? (size_t)q.n_ref * q.d * esz : 0);
            q.query = base;
            q.ld_query = q.d;
            if (q.ref) {
                q.ref = p_ref;
                q.ld_ref = q.d;
            }
            q.logl = reinterpret_cast<const double*>(p_logl);
            q.weights = reinterpret_cast<const double*>(p_logl + (size_t)q.n_query * 8);
        }
        unsigned hw = std::thread::hardware_concurrency();
        const int nthreads = (int)std::min<unsigned>(hw ? hw : 4u, std::min<unsigned>(8u, (unsigned)n));
        for (int t = 0; t < nthreads; ++t) {
            packers.emplace_back([&, pack_rows]() {
                for (;;) {
                    const int j = next_job.fetch_add(1);
                    if (j >= n) return;
                    const int i = order[j];
                    const mce_problem& src = problems[i];
                    const mce_problem& dst = staged[i];
                    const size_t esz = (src.flags & MCE_FLAG_F32_SAMPLES) ? 4 : 8;
                    pack_rows(static_cast<unsigned char*>(const_cast<void*>(dst.query)), src.query, src.n_query, src.ld_query, src.d, esz);
                    if (src.ref)
                        pack_rows(static_cast<unsigned char*>(const_cast<void*>(dst.ref)), src.ref, src.n_ref, src.ld_ref, src.d, esz);
                    memcpy(const_cast<double*>(dst.logl), src.logl, (size_t)src.n_query * 8);
                    memcpy(const_cast<double*>(dst.weights), src.weights, (size_t)src.n_query * 8);
                    ready[i].store(1, std::memory_order_release);
                }
            });
        }
    }
    std::vector<Pending> pend(n);
    int failed = -1, failed_rc = 0;
    for (int j = 0; j < n && failed < 0; ++j) {
        const int i = order[j];
        mce_ctx* lane = c->lanes[j % nl];
        if (stage)
            while (ready[i].load(std::memory_order_acquire) == 0) std::this_thread::yield();
        Run r{stage ? &staged[i] : &problems[i], true, 0, nullptr, nullptr, &results[i], nullptr};
        r.hres = c->hres.as<DevScalars>() + i;
        r.batch_lane = true;
        pend[i].ev0 = c->ev_pool[2 * i];
        pend[i].ev1 = c->ev_pool[2 * i + 1];
        rc = enqueue_problem(lane, r, pend[i]);
        if (rc) {
            failed = i;
            failed_rc = rc;
            named(i, rc);
        }
    }
    char first_err[sizeof(g_err)];
    snprintf(first_err, sizeof(first_err), "%s", g_err);
    if (failed >= 0) next_job.store(n);      // nothing more to pack
    for (auto& t : packers) t.join();
    for (int l = 0; l < nl; ++l) {
        const cudaError_t e = cudaStreamSynchronize(c->lanes[l]->stream);
        if (e != cudaSuccess && failed < 0) {
            failed = n;
            failed_rc = MCE_ERR_CUDA;
            snprintf(first_err, sizeof(first_err), "batch lane %d: %s", l, cudaGetErrorString(e));
        }
    }
    cudaGetLastError();
    if (failed >= 0) return fail(failed_rc, "%s", first_err);
    for (int j = 0; j < n; ++j) {      // in enqueue order: the first failing problem is reported
        const int i = order[j];
        rc = complete_problem(c->lanes[j % nl], pend[i]);
        if (rc) return named(i, rc);
    }
    return MCE_OK;
}

// Batch-size sweep (MCEvidence.py:808-840, 1039-1045): the evidence of the first batch_rows[b] rows of the query set for
// every b, from one upload, one whitening and ONE pass of the filter over the largest batch.
int mce_evidence_sweep(mce_ctx* c, const mce_problem* problem, int nbatch, const int64_t* batch_rows, mce_result* results) {
    if (!c) return fail(MCE_ERR_ARG, "ctx is NULL");
    if (!problem || nbatch < 1 || !batch_rows || !results) return fail(MCE_ERR_ARG, "bad sweep arguments");
    std::lock_guard<std::mutex> lk(c->mu);
    Sweep sw;
    sw.nb = nbatch;
    sw.rows = batch_rows;
    sw.results = results;
    Run r{problem, true, 0, nullptr, nullptr, &results[0], nullptr};
    r.sweep = &sw;
    return run_problem(c, r);
}

int mce_knn(mce_ctx* c, const mce_problem* problem, int n_neighbors, double* dknn, int64_t* idx, mce_stats* stats) {
    if (!c) return fail(MCE_ERR_ARG, "ctx is NULL");
    std::lock_guard<std::mutex> lk(c->mu);
    Run r{problem, false, n_neighbors, dknn, idx, nullptr, stats};
    return run_problem(c, r);
}

int mce_combine_partials(int nparts, int nk, const double* parts, double* out) {
    if (nparts < 1 || nk < 1 || !parts || !out) return fail(MCE_ERR_ARG, "bad arguments to mce_combine_partials");
    for (int k = 0; k < nk; ++k) {
        double m = -INFINITY;
        for (int i = 0; i < nparts; ++i) {
            const double v = parts[((size_t)i * nk + k) * 3];
            if (v > m || isnan(v)) m = v;
        }
        double s1 = 0.0, s2 = 0.0;
        for (int i = 0; i < nparts; ++i) {
            const double* t = parts + ((size_t)i * nk + k) * 3;
            if (t[0] == -INFINITY) continue;
            const double f = exp(t[0] - m);
            s1 += t[1] * f;
            s2 += t[2] * f * f;
        }
        out[k * 3] = m;
        out[k * 3 + 1] = s1;
        out[k * 3 + 2] = s2;
    }
    return MCE_OK;
}

int mce_finalize(const double* partials, int nk, int d, int cross, int64_t n_query_total, double log_sum_w,
                 double log_jacobian, double log_prior_volume, double* logE, double* var_logE) {
    if (!partials || !logE || nk < 1 || d < 1) return fail(MCE_ERR_ARG, "bad arguments to mce_finalize");
    const double logc = 0.5 * d * log(M_PI) - lgamma(1.0 + 0.5 * d);  // log(pi^{d/2} / Gamma(1 + d/2)), MCEvidence.py:1110
    for (int k = 1; k <= nk; ++k) {
        const double* t = partials + (size_t)(k - 1) * 3;
        const int k_nn = cross ? k + 1 : k;                              // MCEvidence.py:1120-1122
        logE[k - 1] = log_sum_w + log_jacobian - log_prior_volume - log((double)n_query_total * k_nn + 1.0) + logc +
                      t[0] + log(t[1]);                                  // MCEvidence.py:1123-1131
        if (var_logE) var_logE[k - 1] = t[2] / (t[1] * t[1]) - 1.0 / (double)n_query_total;
    }
    return MCE_OK;
}

int mce_plan_split(int sm_count, int d, int64_t n_query, int64_t n_ref, int list_len, int queries_per_cta,
                   int ctas_per_sm, int wide, int* nfull, int* chunks, double* cost) {
    if (sm_count < 1 || d < 1 || n_query < 1 || n_ref < 1 || list_len < 1 || queries_per_cta < 1 || ctas_per_sm < 1 ||
        !nfull || !chunks)
        return fail(MCE_ERR_ARG, "bad arguments to mce_plan_split");
    const int TJ = tile_refs(d);
    const int ntiles = (int)((n_ref + TJ - 1) / TJ);
    const int max_chunks = ntiles / 4 > 1 ? (ntiles / 4 > 64 ? 64 : ntiles / 4) : 1;
    const SplitPlan pl = plan_split((n_query + queries_per_cta - 1) / queries_per_cta, sm_count, queries_per_cta,
                                    ctas_per_sm, n_ref, list_len, d, wide != 0, max_chunks);
    *nfull = pl.nfull;
    *chunks = pl.chunks;
    if (cost) *cost = pl.cost;
    return MCE_OK;
}

int mce_fma_peak(mce_ctx* c, int variant, int iters, double* tflops, double* ms_out) {
    if (!c || !tflops) return fail(MCE_ERR_ARG, "ctx or tflops is NULL");
    if (variant < 0 || variant > 4 || iters < 1) return fail(MCE_ERR_ARG, "bad variant/iters");
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    const int grid = c->sm_count * 8;
    int rc = c->sink.ensure((size_t)grid * 256 * sizeof(float));
    if (rc) return rc;
    cudaStream_t st = c->stream;
    launch_fma_peak(variant, iters / 8 + 1, grid, c->sink.as<float>(), st);  // warm-up
    CU(cudaEventRecord(c->ev[EV_START], st));
    launch_fma_peak(variant, iters, grid, c->sink.as<float>(), st);
    CU(cudaEventRecord(c->ev[EV_END], st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, c->ev[EV_START], c->ev[EV_END]));
    *tflops = (double)fma_peak_flops(variant, iters, grid) / ((double)ms * 1e-3) / 1e12;
    if (ms_out) *ms_out = ms;
    return MCE_OK;
}

int mce_unit_peak(mce_ctx* c, int unit, int iters, double* rate, double* ms_out) {
    if (!c || !rate) return fail(MCE_ERR_ARG, "ctx or rate is NULL");
    if (unit < 0 || unit > 3 || iters < 1) return fail(MCE_ERR_ARG, "bad unit/iters");
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    int rc = c->sink.ensure((size_t)c->sm_count * 8 * 256 * sizeof(float));
    if (rc) return rc;
    cudaStream_t st = c->stream;
    int grid = 0;
    if (launch_unit_peak(unit, iters / 8 + 1, c->sm_count, c->sink.as<float>(), st, &grid) != 0)   // warm-up
        return fail(MCE_ERR_CUDA, "unit-peak launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    CU(cudaEventRecord(c->ev[EV_START], st));
    if (launch_unit_peak(unit, iters, c->sm_count, c->sink.as<float>(), st, &grid) != 0)
        return fail(MCE_ERR_CUDA, "unit-peak launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    CU(cudaEventRecord(c->ev[EV_END], st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, c->ev[EV_START], c->ev[EV_END]));
    *rate = unit_peak_work(unit, iters, grid) / ((double)ms * 1e-3);
    if (ms_out) *ms_out = ms;
    return MCE_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// Device group: one box, one process, one host thread per GPU.
// NCCL is bound at run time (dlopen of libnccl.so.2: the library links against nothing but the CUDA runtime, and a
// Python process keeps whatever NCCL torch brought); without it every device uploads the host arrays itself.
namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load() {
        if (handle) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            handle = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
            if (handle) break;
        }
        if (!handle) return false;
        CommInitAll = reinterpret_cast<decltype(CommInitAll)>(dlsym(handle, "ncclCommInitAll"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(handle, "ncclCommDestroy"));
        Broadcast = reinterpret_cast<decltype(Broadcast)>(dlsym(handle, "ncclBroadcast"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(handle, "ncclGetErrorString"));
        if (!CommInitAll || !CommDestroy || !Broadcast || !GetErrorString) {
            dlclose(handle);
            handle = nullptr;
            return false;
        }
        return true;
    }
};

}  // namespace

struct mce_group {
    std::vector<mce_ctx*> ctx;          // one per device, rank order
    std::vector<ncclComm_t> comm;       // empty without NCCL
    std::vector<DevBuf> raw_q, raw_r, logl, w;   // per rank: broadcast samples, this rank's slices of logL and weights
    NcclApi nccl;
    std::mutex mu;
};

namespace {

void shard_rows(long long n, int rank, int world, long long* lo, long long* hi) {
    const long long base = n / world, rem = n % world;
    *lo = rank * base + (rank < rem ? rank : rem);
    *hi = *lo + base + (rank < rem ? 1 : 0);
}

// One rank of mce_group_evidence: rows [lo, hi) of the query set against the whole reference set -> partial triples.
int group_rank_evidence(mce_group* g, int r, const mce_problem& p, long long lo, long long hi, double* partials,
                        mce_result* sub) {
    mce_ctx* c = g->ctx[r];
    const int world = (int)g->ctx.size();
    CU(cudaSetDevice(c->device));
    mce_problem q = p;
    q.q_begin = lo;
    q.q_end = hi;
    q.flags |= MCE_FLAG_PARTIALS;
    memset(sub, 0, sizeof(*sub));
    sub->partials = partials;
    const bool bcast = !g->comm.empty() && world > 1 && !p.query_rows && !p.ref_rows;
    if (bcast) {
        // the sample arrays go to rank 0 once and from there over NVLink to everybody (ncclBroadcast on every rank's
        // own stream, which also orders it before the kernels); logL and the weights are needed for this rank's rows only
        cudaStream_t st = c->stream;
        const size_t esz = (p.flags & MCE_FLAG_F32_SAMPLES) ? sizeof(float) : sizeof(double);
        auto share = [&](DevBuf& buf, const void* host, long long n, long long ld) -> int {
            const size_t bytes = ((size_t)(n - 1) * ld + p.d) * esz;
            int rc = buf.ensure(bytes);
            if (rc) return rc;
            if (r == 0) CU(cudaMemcpyAsync(buf.p, host, bytes, cudaMemcpyHostToDevice, st));
            const ncclResult_t nr = g->nccl.Broadcast(buf.p, buf.p, bytes, ncclChar, 0, g->comm[r], st);
            if (nr != ncclSuccess) return fail(MCE_ERR_CUDA, "ncclBroadcast failed on rank %d: %s", r, g->nccl.GetErrorString(nr));
            return MCE_OK;
        };
        int rc = share(g->raw_q[r], p.query, p.n_query, p.ld_query);
        if (rc) return rc;
        q.query = g->raw_q[r].p;
        if (p.ref) {
            rc = share(g->raw_r[r], p.ref, p.n_ref, p.ld_ref);
            if (rc) return rc;
            q.ref = g->raw_r[r].p;
        }
        const size_t nb = (size_t)(hi - lo) * 8;
        rc = g->logl[r].ensure(nb);
        if (rc) return rc;
        rc = g->w[r].ensure(nb);
        if (rc) return rc;
        CU(cudaMemcpyAsync(g->logl[r].p, p.logl + lo, nb, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(g->w[r].p, p.weights + lo, nb, cudaMemcpyHostToDevice, st));
        q.logl = g->logl[r].as<double>() - lo;      // indexed by global row
        q.weights = g->w[r].as<double>() - lo;
        q.flags |= MCE_FLAG_DEVICE_PTRS;
    }
    return mce_evidence(c, &q, sub);
}

}  // namespace

extern "C" {

int mce_group_create(int n_devices, const int* device_ids, mce_group** out) {
    if (!out) return fail(MCE_ERR_ARG, "out is NULL");
    *out = nullptr;
    const int have = mce_device_count();
    if (have == 0) return fail(MCE_ERR_CUDA, "no CUDA device available; mcevidence_b200 has no CPU fallback");
    if (n_devices <= 0) n_devices = have;
    if (n_devices > have) return fail(MCE_ERR_ARG, "%d devices requested, %d present", n_devices, have);
    mce_group* g = new (std::nothrow) mce_group();
    if (!g) return fail(MCE_ERR_NOMEM, "out of host memory");
    std::vector<int> ids(n_devices);
    for (int i = 0; i < n_devices; ++i) ids[i] = device_ids ? device_ids[i] : i;
    for (int i = 0; i < n_devices; ++i) {
        mce_ctx* c = nullptr;
        const int rc = mce_ctx_create(ids[i], &c);
        if (rc) {
            mce_group_destroy(g);
            return rc;
        }
        g->ctx.push_back(c);
    }
    g->raw_q.resize(n_devices);
    g->raw_r.resize(n_devices);
    g->logl.resize(n_devices);
    g->w.resize(n_devices);
    const char* no_nccl = getenv("MCE_GROUP_NO_NCCL");
    if (n_devices > 1 && !(no_nccl && no_nccl[0] == '1') && g->nccl.load()) {
        g->comm.resize(n_devices);
        const ncclResult_t nr = g->nccl.CommInitAll(g->comm.data(), n_devices, ids.data());
        if (nr != ncclSuccess) {
            g->comm.clear();     // fall back to one upload per device
            cudaGetLastError();
        }
    }
    *out = g;
    return MCE_OK;
}

void mce_group_destroy(mce_group* g) {
    if (!g) return;
    for (size_t r = 0; r < g->ctx.size(); ++r) {
        cudaSetDevice(g->ctx[r]->device);
        if (r < g->comm.size()) g->nccl.CommDestroy(g->comm[r]);
        if (r < g->raw_q.size()) {
            g->raw_q[r].release();
            g->raw_r[r].release();
            g->logl[r].release();
            g->w[r].release();
        }
        mce_ctx_destroy(g->ctx[r]);
    }
    delete g;
}

int mce_group_size(const mce_group* g) { return g ? (int)g->ctx.size() : 0; }
int mce_group_uses_nccl(const mce_group* g) { return g && !g->comm.empty() ? 1 : 0; }

int mce_group_set_option(mce_group* g, const char* name, int value) {
    if (!g) return fail(MCE_ERR_ARG, "group is NULL");
    for (mce_ctx* c : g->ctx) {
        const int rc = mce_ctx_set_option(c, name, value);
        if (rc) return rc;
    }
    return MCE_OK;
}

int mce_group_evidence(mce_group* g, const mce_problem* problem, mce_result* result) {
    if (!g || !problem || !result) return fail(MCE_ERR_ARG, "group, problem or result is NULL");
    if (!result->logE) return fail(MCE_ERR_ARG, "result.logE is NULL");
    if (problem->flags & MCE_FLAG_DEVICE_PTRS) return fail(MCE_ERR_ARG, "a device group takes host arrays");
    if (problem->q_begin != 0 || problem->q_end != 0) return fail(MCE_ERR_ARG, "a device group shards the rows itself: no query row range");
    if (result->dknn || result->idx) return fail(MCE_ERR_ARG, "neighbour arrays are not gathered by a device group");
    int rc = validate(problem, true);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(g->mu);
    const int world = (int)g->ctx.size();
    const int kmax = problem->kmax < 2 ? 2 : problem->kmax;
    const int nk = kmax - 1;
    std::vector<double> parts((size_t)world * nk * 3);
    std::vector<mce_result> subs(world);
    std::vector<int> rcs(world, MCE_OK);
    std::vector<std::string> errs(world);
    std::vector<std::thread> threads;
    for (int r = 0; r < world; ++r) {
        threads.emplace_back([&, r]() {
            long long lo, hi;
            shard_rows(problem->n_query, r, world, &lo, &hi);
            if (hi <= lo) {          // more devices than rows: an empty shard contributes nothing
                for (int k = 0; k < nk; ++k) {
                    parts[((size_t)r * nk + k) * 3] = -INFINITY;
                    parts[((size_t)r * nk + k) * 3 + 1] = 0.0;
                    parts[((size_t)r * nk + k) * 3 + 2] = 0.0;
                }
                memset(&subs[r], 0, sizeof(mce_result));
                return;
            }
            rcs[r] = group_rank_evidence(g, r, *problem, lo, hi, parts.data() + (size_t)r * nk * 3, &subs[r]);
            if (rcs[r]) errs[r] = g_err;     // the message is thread-local
        });
    }
    for (auto& t : threads) t.join();
    for (int r = 0; r < world; ++r)
        if (rcs[r]) return fail(rcs[r], "device %d: %s", g->ctx[r]->device, errs[r].c_str());
    // the exchange after the search is 3 (kmax - 1) doubles per device, already on the host of this one process
    std::vector<double> total((size_t)nk * 3);
    rc = mce_combine_partials(world, nk, parts.data(), total.data());
    if (rc) return rc;
    result->log_jacobian = subs[0].log_jacobian;
    result->stats = subs[0].stats;
    for (int r = 1; r < world; ++r) {     // the slowest device sets the time; counts add up
        if (subs[r].stats.ms_total > result->stats.ms_total) result->stats.ms_total = subs[r].stats.ms_total;
        result->stats.pairs += subs[r].stats.pairs;
        result->stats.rows_rescanned += subs[r].stats.rows_rescanned;
        result->stats.launches += subs[r].stats.launches;
    }
    if ((problem->flags & MCE_FLAG_PARTIALS) && result->partials) memcpy(result->partials, total.data(), total.size() * 8);
    return mce_finalize(total.data(), nk, problem->d, problem->ref ? 1 : 0, problem->n_query, problem->log_sum_w,
                        result->log_jacobian, problem->log_prior_volume, result->logE, result->var_logE);
}

int mce_group_evidence_batch(mce_group* g, int n, const mce_problem* problems, mce_result* results) {
    if (!g) return fail(MCE_ERR_ARG, "group is NULL");
    if (n < 0 || (n > 0 && (!problems || !results))) return fail(MCE_ERR_ARG, "bad batch arguments");
    std::lock_guard<std::mutex> lk(g->mu);
    const int world = (int)g->ctx.size();
    // chain-per-GPU: largest first to the least loaded device (planck_mcevidence.py:149-156 splits per data set)
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    auto cost = [&](int i) { return (double)problems[i].n_query * (double)(problems[i].ref ? problems[i].n_ref : problems[i].n_query); };
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost(a) > cost(b); });
    std::vector<std::vector<int>> mine(world);
    std::vector<double> load(world, 0.0);
    for (int i : order) {
        const int r = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        mine[r].push_back(i);
        load[r] += cost(i);
    }
    std::vector<int> rcs(world, MCE_OK);
    std::vector<std::string> errs(world);
    std::vector<std::thread> threads;
    for (int r = 0; r < world; ++r) {
        threads.emplace_back([&, r]() {
            const int m = (int)mine[r].size();
            if (m == 0) return;
            std::vector<mce_problem> ps(m);
            std::vector<mce_result> rs(m);
            for (int j = 0; j < m; ++j) {
                ps[j] = problems[mine[r][j]];
                rs[j] = results[mine[r][j]];
            }
            rcs[r] = mce_evidence_batch(g->ctx[r], m, ps.data(), rs.data());
            if (rcs[r]) errs[r] = g_err;
            for (int j = 0; j < m; ++j) results[mine[r][j]] = rs[j];
        });
    }
    for (auto& t : threads) t.join();
    for (int r = 0; r < world; ++r)
        if (rcs[r]) return fail(rcs[r], "device %d: %s", g->ctx[r]->device, errs[r].c_str());
    return MCE_OK;
}

}  // extern "C"
