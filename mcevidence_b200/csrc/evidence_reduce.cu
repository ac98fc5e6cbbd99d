// K4: fused log-volume + logsumexp epilogue.
//
// Restates the volume loop and the reduction of MCEvidence.py:1107-1131 in log space:
//   volume_jk = c_d r_jk^d,  dotp_k = sum_j volume_jk / w_j * exp(logL_j - logLmax)
//   t_jk = d log r_jk + logL_j - log w_j   =>   log dotp_k = log c_d - logLmax + logsumexp_j t_jk
// (logLmax cancels in MLE, MCEvidence.py:1131).  Per output k the kernel produces the triple
// (M, S1, S2) = (max_j t, sum exp(t - M), sum exp(2 (t - M))); S2 feeds the delta-method
// variance, which has no counterpart in the reference.  Rows with r = 0 (duplicates) have
// t = -inf and contribute exactly 0, as volume = 0 does at MCEvidence.py:1110.
// Summation order is fixed (grid size depends only on the row count): results are
// reproducible run to run.
#include "../../include/mcevidence_b200.h"
#include "common.cuh"
#include <math_constants.h>

namespace mce {

namespace {

constexpr int RED_THREADS = 256;

struct Lse {
    double m, s1, s2;
    __device__ __forceinline__ void init() { m = -CUDART_INF; s1 = 0.0; s2 = 0.0; }
    __device__ __forceinline__ void add(double t) {
        if (t == -CUDART_INF) return;
        if (t > m) {
            if (m != -CUDART_INF) {
                const double f = exp(m - t);
                s1 *= f;
                s2 *= f * f;
            }
            m = t;
            s1 += 1.0;
            s2 += 1.0;
        } else {
            const double e = exp(t - m);  // NaN propagates; t = +inf makes m = +inf
            s1 += e;
            s2 += e * e;
        }
    }
    __device__ __forceinline__ void merge(double om, double os1, double os2) {
        if (om == -CUDART_INF) return;
        if (m == -CUDART_INF) { m = om; s1 = os1; s2 = os2; return; }
        if (om > m) {
            const double f = exp(m - om);
            s1 = s1 * f + os1;
            s2 = s2 * f * f + os2;
            m = om;
        } else {
            const double f = exp(om - m);
            s1 += os1 * f;
            s2 += os2 * f * f;
        }
    }
};

constexpr int MAX_NK = MCE_MAX_NEIGHBOURS;   // kmax - 1 <= 16 outputs

// warp-level merge (fixed shuffle tree: deterministic)
__device__ __forceinline__ void warp_merge(Lse& v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double om = __shfl_down_sync(0xffffffffu, v.m, o);
        const double o1 = __shfl_down_sync(0xffffffffu, v.s1, o);
        const double o2 = __shfl_down_sync(0xffffffffu, v.s2, o);
        v.merge(om, o1, o2);
    }
}

// One pass over the rows: logL_j - log w_j is evaluated once per row and shared by all k.
__global__ void __launch_bounds__(RED_THREADS) partials_kernel(const ReduceParams p) {
    __shared__ double sh[MAX_NK][RED_THREADS / 32][3];
    const long long nq = p.q_end - p.q_begin;
    const double dd = (double)p.d;
    const int col0 = p.col0 + p.first_nb - 1;
    Lse acc[MAX_NK];
#pragma unroll
    for (int k = 0; k < MAX_NK; ++k) acc[k].init();
    for (long long ql = (long long)blockIdx.x * RED_THREADS + threadIdx.x; ql < nq;
         ql += (long long)gridDim.x * RED_THREADS) {
        const long long row = p.q_begin + ql;
        const double base = p.logl[row] - log(p.weights[row]);
        const double* r = p.dknn + ql * p.ld + col0;
#pragma unroll
        for (int k = 0; k < MAX_NK; ++k)
            if (k < p.nk) acc[k].add(dd * log(r[k]) + base);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < MAX_NK; ++k) {
        if (k < p.nk) {
            warp_merge(acc[k]);
            if (lane == 0) { sh[k][warp][0] = acc[k].m; sh[k][warp][1] = acc[k].s1; sh[k][warp][2] = acc[k].s2; }
        }
    }
    __syncthreads();
    if (threadIdx.x < p.nk) {
        const int k = threadIdx.x;
        Lse v;
        v.init();
        for (int w = 0; w < RED_THREADS / 32; ++w) v.merge(sh[k][w][0], sh[k][w][1], sh[k][w][2]);
        double* o = p.block_partial + ((long long)blockIdx.x * p.nk + k) * 3;
        o[0] = v.m; o[1] = v.s1; o[2] = v.s2;
    }
}

// One warp per output k: lanes stride over the block partials, then a fixed shuffle tree.
__global__ void final_kernel(const double* block_partial, int nblocks, int nk, double* out) {
    const int k = blockIdx.x;
    const int lane = threadIdx.x;
    Lse acc;
    acc.init();
    for (int b = lane; b < nblocks; b += 32) {
        const double* o = block_partial + ((long long)b * nk + k) * 3;
        acc.merge(o[0], o[1], o[2]);
    }
    warp_merge(acc);
    if (lane == 0) {
        out[k * 3] = acc.m;
        out[k * 3 + 1] = acc.s1;
        out[k * 3 + 2] = acc.s2;
    }
}

}  // namespace

int reduce_grid_size(long long nq) {
    long long g = (nq + RED_THREADS * 8 - 1) / (RED_THREADS * 8);
    if (g < 1) g = 1;
    if (g > 1024) g = 1024;
    return (int)g;
}

void launch_reduce(const ReduceParams& p, int grid, cudaStream_t st, int* launches) {
    partials_kernel<<<grid, RED_THREADS, 0, st>>>(p);
    final_kernel<<<p.nk, 32, 0, st>>>(p.block_partial, grid, p.nk, p.out);
    *launches += 2;
}

}  // namespace mce
