// K3: FP64 re-rank of the filter's candidates, per-row exactness verification, and the exact
// FP64 re-scan of the rows that cannot be verified.
//
// Output contract = NearestNeighbors.kneighbors at MCEvidence.py:1104: ascending Euclidean
// distances in float64 plus int64 indices; ties are ordered by index (sklearn leaves the order
// of exact ties unspecified).  Squared distances are summed dimension by dimension without FMA
// contraction, the arithmetic of sklearn's kd-tree metric (metrics/_dist_metrics.pxd.tp:39-49).
//
// Verification.  The filter ranks by the FP32 score s = fl(q.r - |r|^2/2).  For every reference
// the filter did NOT keep in chunk c, fl(s) <= th_c (the chunk's final threshold).  With
//   eps = u * ((d+2) |q| Rmax + (d+1)/2 Rmax^2) * 1.0625,  u = 2^-24
// bounding the FP32 rounding of the inputs, of h and of the d chained FMAs, the true squared
// distance of every such reference is >= |q|^2 - 2 (max_c th_c + eps) =: B.  If the k-th exact
// distance among the candidates is < B, the k nearest neighbours (and their order, ties
// included) are exact.  Otherwise the row is re-scanned exhaustively in FP64.
#include "common.cuh"
#include "knn_filter.cuh"
#include <math_constants.h>

namespace mce {

int filter_query(int d, const FilterConfig& cfg, int list_len, int* qpc, int* smem, int* ctas) {
    if (d >= 1 && d <= 16) return filter_dispatch_lo(0, d, cfg, nullptr, list_len, 0, qpc, smem, ctas);
    if (d <= 32) return filter_dispatch_mid(0, d, cfg, nullptr, list_len, 0, qpc, smem, ctas);
    if (d <= MAX_DIM) return filter_dispatch_hi(0, d, cfg, nullptr, list_len, 0, qpc, smem, ctas);
    return -1;
}

int launch_filter(const FilterParams& p, const FilterConfig& cfg, cudaStream_t st, int* launches) {
    int rc = -1;
    if (p.d >= 1 && p.d <= 16) rc = filter_dispatch_lo(1, p.d, cfg, &p, p.list_len, st, nullptr, nullptr, nullptr);
    else if (p.d <= 32) rc = filter_dispatch_mid(1, p.d, cfg, &p, p.list_len, st, nullptr, nullptr, nullptr);
    else if (p.d <= MAX_DIM) rc = filter_dispatch_hi(1, p.d, cfg, &p, p.list_len, st, nullptr, nullptr, nullptr);
    if (rc == 0) *launches += 1;
    return rc;
}

void filter_default_config(int d, int list_len, FilterConfig* cfg) {
    cfg->packed = 1;
    // measured on B200 (profiles/r01_filter_variants.md): every tiling saturates near 55-60 TFLOP/s
    // of executed FMA; two queries per thread x eight references per step with a CTA barrier per
    // tile group and a one-row register prefetch is within 2 % of the best at d = 10, 20 and 27
    if (d <= 32) { cfg->tq = 2; cfg->tr = 8; }
    else { cfg->tq = 2; cfg->tr = 4; }
    cfg->sync = 2;
    (void)list_len;
}

namespace {

__device__ __forceinline__ bool lex_less(double da, long long ia, double db, long long ib) {
    return da < db || (da == db && ia < ib);
}

__device__ __forceinline__ double sq_dist_seq(const double* __restrict__ a, const double* __restrict__ b, int d) {
    double acc = 0.0;
    for (int c = 0; c < d; ++c) {
        const double t = __dsub_rn(a[c], b[c]);
        acc = __dadd_rn(acc, __dmul_rn(t, t));
    }
    return acc;
}

// sorted (ascending, lexicographic) fixed-capacity list in registers
struct TopList {
    double dist[KCAP];
    long long idx[KCAP];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int i = 0; i < KCAP; ++i) {
            dist[i] = CUDART_INF;
            idx[i] = 0x7fffffffffffffffLL;
        }
    }
    __device__ __forceinline__ void push(double d2, long long id) {
        if (!lex_less(d2, id, dist[KCAP - 1], idx[KCAP - 1])) return;
        dist[KCAP - 1] = d2;
        idx[KCAP - 1] = id;
#pragma unroll
        for (int i = KCAP - 1; i > 0; --i) {
            if (lex_less(dist[i], idx[i], dist[i - 1], idx[i - 1])) {
                const double td = dist[i]; dist[i] = dist[i - 1]; dist[i - 1] = td;
                const long long ti = idx[i]; idx[i] = idx[i - 1]; idx[i - 1] = ti;
            }
        }
    }
    __device__ __forceinline__ void pop_front() {
#pragma unroll
        for (int i = 0; i < KCAP - 1; ++i) {
            dist[i] = dist[i + 1];
            idx[i] = idx[i + 1];
        }
        dist[KCAP - 1] = CUDART_INF;
        idx[KCAP - 1] = 0x7fffffffffffffffLL;
    }
};

// `row` = output row, `self` = reference index written into the synthetic self column
__device__ __forceinline__ void write_row(const RerankParams& p, long long row, long long self, const TopList& top) {
    double* drow = p.dknn + row * p.out_ld;
    long long* irow = p.idx ? p.idx + row * p.out_ld : nullptr;
    if (p.write_self_col) {
        drow[0] = 0.0;
        if (irow) irow[0] = self;
    }
#pragma unroll
    for (int i = 0; i < KCAP; ++i) {
        if (i < p.k) {
            drow[p.out_col0 + i] = sqrt(top.dist[i]);
            if (irow) irow[p.out_col0 + i] = top.dist[i] < CUDART_INF ? top.idx[i] : -1;
        }
    }
}

__global__ void __launch_bounds__(128) rerank_kernel(const RerankParams p) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool sweep = p.sweep_nb > 0;
    const long long nrows = sweep ? p.sweep_off[p.sweep_nb] : p.q_end - p.q_begin;
    if (t >= nrows || *p.status != 0) return;
    if (p.flag_all) {
        p.flagged[t] = (int)t;
        if (t == 0) *p.nflagged = (int)nrows;
        return;
    }
    // (batch, query) of this row; a single problem is batch 0 with every chunk
    long long ql = t;
    int nchunks;
    long long edge = 0, edge_end = 0;
    if (sweep) {
        const int b = sweep_batch_of(p, t);
        ql = t - p.sweep_off[b];
        nchunks = p.sweep_nchunk[b];
        edge = p.sweep_edge[b];
        edge_end = p.sweep_rows[b];
    } else {
        nchunks = ql < p.nfull_queries ? 1 : p.ref_chunks;
    }
    const int d = p.d;
    const double* q = p.yq + (p.q_begin + ql) * d;
    const long long self = p.q_begin + ql + p.self_offset;
    TopList top;
    top.init();
    float thmax = -CUDART_INF_F;
    for (int r = 0; r < nchunks; ++r) thmax = fmaxf(thmax, p.cand_th[cand_th_offset(r, ql, p.nq_pad, p.ntail_pad, p.tail_q0)]);
    // one candidate stream (a single insertion site keeps the sorted list in registers): the filter's lists chunk by
    // chunk, then -- sweep only -- the references of the partial tile behind the last whole chunk
    const long long nlist = (long long)nchunks * p.list_len;
    const long long ncand = nlist + (edge_end - edge);
    int r = 0, e = 0;
    for (long long i = 0; i < ncand; ++i) {
        long long id;
        if (i < nlist) {
            id = p.cand_idx[cand_idx_offset(r, e, ql, p.list_len, p.nq_pad, p.ntail_pad, p.tail_q0)];
            if (++e == p.list_len) {
                e = 0;
                ++r;
            }
            if (id < 0) continue;
        } else {
            id = edge + (i - nlist);
            if (p.self_exclude && id == self) continue;
        }
        top.push(sq_dist_seq(q, p.yr + id * d, d), id);
    }
    double qn2 = 0.0;
    for (int c = 0; c < d; ++c) qn2 += q[c] * q[c];
    const double rmax = (double)__uint_as_float(*p.rmax_bits);
    const double u = 5.9604644775390625e-08;  // 2^-24
    const double eps = 1.0625 * u * ((d + 2) * sqrt(qn2) * rmax + 0.5 * (d + 1) * rmax * rmax);
    const double bound = qn2 - 2.0 * ((double)thmax + eps);
    const bool ok = (thmax == -CUDART_INF_F) || (top.dist[p.k - 1] < bound);
    write_row(p, t, self, top);
    if (!ok) {
        const int pos = atomicAdd(p.nflagged, 1);
        p.flagged[pos] = (int)t;
    }
}

constexpr int EXACT_THREADS = 256;
constexpr int EXACT_MAX_SEGMENTS = 64;

// Few flagged rows: each row is split into `seg` reference segments handled by different CTAs
// (partial lists go to scratch and are merged by exact_merge_kernel).  Many rows: one CTA per row.
__device__ __forceinline__ int exact_segments(int nfl, int grid) {
    if (nfl <= 0) return 1;
    int s = grid / nfl;
    if (s < 1) s = 1;
    if (s > EXACT_MAX_SEGMENTS) s = EXACT_MAX_SEGMENTS;
    return s;
}

// Exhaustive FP64 scan of the flagged rows: per-thread sorted lists, k rounds of block argmin.
__global__ void __launch_bounds__(EXACT_THREADS) exact_rescan_kernel(const RerankParams p, double* scratch_d,
                                                                      long long* scratch_i) {
    __shared__ double sd[EXACT_THREADS / 32];
    __shared__ long long si[EXACT_THREADS / 32];
    __shared__ int sowner[EXACT_THREADS / 32];
    __shared__ double wd;
    __shared__ long long wi;
    __shared__ int wowner;
    __shared__ double sq[MAX_DIM];
    const int nfl = *p.status != 0 ? 0 : *p.nflagged;
    const int d = p.d;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nseg = exact_segments(nfl, gridDim.x);
    for (long long item = blockIdx.x; item < (long long)nfl * nseg; item += gridDim.x) {
        const int f = (int)(item / nseg);
        const int seg = (int)(item % nseg);
        const long long row = p.flagged[f];
        long long ql = row, n_ref = p.n_ref;
        if (p.sweep_nb > 0) {
            const int b = sweep_batch_of(p, row);
            ql = row - p.sweep_off[b];
            n_ref = p.sweep_rows[b];
        }
        const long long seg_len = (n_ref + nseg - 1) / nseg;
        const double* qg = p.yq + (p.q_begin + ql) * d;
        __syncthreads();
        for (int c = threadIdx.x; c < d; c += EXACT_THREADS) sq[c] = qg[c];
        __syncthreads();
        const long long self = p.self_exclude ? p.q_begin + ql + p.self_offset : -1;
        const long long j_lo = seg * seg_len;
        const long long j_hi = min(n_ref, j_lo + seg_len);
        TopList top;
        top.init();
        for (long long j = j_lo + threadIdx.x; j < j_hi; j += EXACT_THREADS) {
            if (j == self) continue;
            top.push(sq_dist_seq(sq, p.yr + j * d, d), j);
        }
        TopList res;
        res.init();
        for (int round = 0; round < p.k; ++round) {
            double bd = top.dist[0];
            long long bi = top.idx[0];
            int bo = threadIdx.x;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double od = __shfl_xor_sync(0xffffffffu, bd, o);
                const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
                const int oo = __shfl_xor_sync(0xffffffffu, bo, o);
                if (lex_less(od, oi, bd, bi)) { bd = od; bi = oi; bo = oo; }
            }
            if (lane == 0) { sd[warp] = bd; si[warp] = bi; sowner[warp] = bo; }
            __syncthreads();
            if (threadIdx.x == 0) {
                double md = sd[0]; long long mi = si[0]; int mo = sowner[0];
                for (int w = 1; w < EXACT_THREADS / 32; ++w)
                    if (lex_less(sd[w], si[w], md, mi)) { md = sd[w]; mi = si[w]; mo = sowner[w]; }
                wd = md; wi = mi; wowner = mo;
            }
            __syncthreads();
            if (threadIdx.x == wowner) top.pop_front();
            if (threadIdx.x == 0) {
#pragma unroll
                for (int i = 0; i < KCAP; ++i)
                    if (i == round) { res.dist[i] = wd; res.idx[i] = wi; }
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            if (nseg == 1) {
                write_row(p, row, p.q_begin + ql + p.self_offset, res);
            } else {
#pragma unroll
                for (int i = 0; i < KCAP; ++i) {
                    scratch_d[item * KCAP + i] = res.dist[i];
                    scratch_i[item * KCAP + i] = res.idx[i];
                }
            }
        }
    }
}

// Merge the per-segment lists of each flagged row (no-op when rows were not segmented).
__global__ void __launch_bounds__(128) exact_merge_kernel(const RerankParams p, int rescan_grid, const double* scratch_d,
                                                           const long long* scratch_i) {
    const int nfl = *p.status != 0 ? 0 : *p.nflagged;
    const int nseg = exact_segments(nfl, rescan_grid);
    if (nseg == 1) return;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nfl) return;
    TopList top;
    top.init();
    for (int seg = 0; seg < nseg; ++seg) {
        const long long item = (long long)f * nseg + seg;
        for (int i = 0; i < p.k; ++i) {
            const double dd = scratch_d[item * KCAP + i];
            if (dd < CUDART_INF) top.push(dd, scratch_i[item * KCAP + i]);
        }
    }
    const long long row = p.flagged[f];
    const long long ql = p.sweep_nb > 0 ? row - p.sweep_off[sweep_batch_of(p, row)] : row;
    write_row(p, row, p.q_begin + ql + p.self_offset, top);
}

}  // namespace

void launch_rerank(const RerankParams& p, cudaStream_t st, int* launches) {
    const long long nq = p.sweep_nb > 0 ? p.sweep_total : p.q_end - p.q_begin;
    if (nq <= 0) return;
    const int threads = 128;
    rerank_kernel<<<(unsigned)((nq + threads - 1) / threads), threads, 0, st>>>(p);
    *launches += 1;
}

size_t exact_scratch_bytes(int grid) { return (size_t)grid * KCAP * (sizeof(double) + sizeof(long long)); }

void launch_exact_rescan(const RerankParams& p, int grid, void* scratch, cudaStream_t st, int* launches) {
    double* sd = reinterpret_cast<double*>(scratch);
    long long* si = reinterpret_cast<long long*>(sd + (size_t)grid * KCAP);
    exact_rescan_kernel<<<grid, EXACT_THREADS, 0, st>>>(p, sd, si);
    // rows are segmented only when fewer than `grid` are flagged, so `grid` threads cover them all
    exact_merge_kernel<<<(grid + 127) / 128, 128, 0, st>>>(p, grid, sd, si);
    *launches += 2;
}

}  // namespace mce
