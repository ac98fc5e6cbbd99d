// K0: unweighted sample covariance (ddof=1) of query U ref, Cholesky factor, log-Jacobian.
// K1: whitening on load, emitting the FP64 rows and the FP32 TMA tiles of the filter.
//
// Restates get_covariance / diagonalise_chain (MCEvidence.py:851-882, 842-849).  The
// reference whitens with eig(C): y = x V / sqrt(lambda).  Any factor W with W W^T = C^{-1}
// gives the same Euclidean distances (they are Mahalanobis distances under C), so the
// Cholesky factor is used: W = L^{-T}, J = sqrt(det C) = prod L_ii.  Rows are centred
// before whitening (the reference does not centre; distances are translation invariant)
// so that the FP32 copy keeps its relative accuracy.
#include "common.cuh"
#include <math.h>

namespace mce {

namespace {

constexpr int COV_THREADS = 256;
constexpr int COV_PAIRS_PER_THREAD = (MAX_DIM * (MAX_DIM + 1) / 2 + COV_THREADS - 1) / COV_THREADS;  // 9

__device__ __forceinline__ double seg_at(const Segment& a, const Segment& b, long long r, int c) {
    return r < a.n ? a.at(r, c) : b.at(r - a.n, c);
}

// Rows per staged chunk so that the chunk fits 32 KB of shared memory.
__host__ __device__ inline int cov_chunk_rows(int d) { return 4096 / d; }

// Stage rows [r0, r0 + nr) of a U b, first d columns (minus `shift` when given), into shared memory, element e = r d + c.
// (row, column) advance incrementally -- no division per element -- and four loads are in flight per thread.
__device__ __forceinline__ void cov_stage_rows(const Segment& a, const Segment& b, long long r0, int nr, int d,
                                               const double* shift, double* xs) {
    const int total = nr * d;
    const int dr = COV_THREADS / d, dc = COV_THREADS - dr * d;   // one stride of COV_THREADS elements = dr rows + dc columns
    int r = threadIdx.x / d, c = threadIdx.x - r * d;
    int e = threadIdx.x;
    for (; e + 3 * COV_THREADS < total; e += 4 * COV_THREADS) {
        int rr[4], cc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            rr[u] = r;
            cc[u] = c;
            r += dr;
            c += dc;
            if (c >= d) {
                c -= d;
                ++r;
            }
        }
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = seg_at(a, b, r0 + rr[u], cc[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) xs[e + u * COV_THREADS] = shift ? v[u] - shift[cc[u]] : v[u];
    }
    for (; e < total; e += COV_THREADS) {
        const double v = seg_at(a, b, r0 + r, c);
        xs[e] = shift ? v - shift[c] : v;
        r += dr;
        c += dc;
        if (c >= d) {
            c -= d;
            ++r;
        }
    }
}

// pass 1: per-CTA column sums
__global__ void __launch_bounds__(COV_THREADS) colsum_kernel(Segment a, Segment b, int d, double* partial) {
    __shared__ double xs[4096];
    const long long n = a.n + b.n;
    const int rc = cov_chunk_rows(d);
    double acc = 0.0;
    for (long long r0 = (long long)blockIdx.x * rc; r0 < n; r0 += (long long)gridDim.x * rc) {
        const int nr = (int)min((long long)rc, n - r0);
        cov_stage_rows(a, b, r0, nr, d, nullptr, xs);
        __syncthreads();
        if (threadIdx.x < d)
            for (int r = 0; r < nr; ++r) acc += xs[r * d + threadIdx.x];
        __syncthreads();
    }
    if (threadIdx.x < d) partial[(long long)blockIdx.x * d + threadIdx.x] = acc;
}

__global__ void mean_kernel(const double* partial, int nparts, int d, long long n, double* mean) {
    const int c = threadIdx.x;
    if (c >= d) return;
    double s = 0.0;
    for (int g = 0; g < nparts; ++g) s += partial[(long long)g * d + c];
    mean[c] = s / (double)n;
}

// pass 2: per-CTA centred second moments, lower triangle packed as p = a(a+1)/2 + b, b <= a
__global__ void __launch_bounds__(COV_THREADS) scatter_kernel(Segment a, Segment b, int d, const double* mean,
                                                               double* partial) {
    __shared__ double xs[4096];
    __shared__ double smean[MAX_DIM];
    const long long n = a.n + b.n;
    const int rc = cov_chunk_rows(d);
    const int npairs = d * (d + 1) / 2;
    if (threadIdx.x < d) smean[threadIdx.x] = mean[threadIdx.x];
    int pa[COV_PAIRS_PER_THREAD], pb[COV_PAIRS_PER_THREAD];
    double acc[COV_PAIRS_PER_THREAD];
#pragma unroll
    for (int i = 0; i < COV_PAIRS_PER_THREAD; ++i) {
        const int p = threadIdx.x + i * COV_THREADS;
        int ia = 0;
        while ((ia + 1) * (ia + 2) / 2 <= p) ++ia;
        pa[i] = ia;
        pb[i] = p - ia * (ia + 1) / 2;
        acc[i] = 0.0;
    }
    __syncthreads();
    for (long long r0 = (long long)blockIdx.x * rc; r0 < n; r0 += (long long)gridDim.x * rc) {
        const int nr = (int)min((long long)rc, n - r0);
        cov_stage_rows(a, b, r0, nr, d, smean, xs);
        __syncthreads();
#pragma unroll
        for (int i = 0; i < COV_PAIRS_PER_THREAD; ++i) {
            if (threadIdx.x + i * COV_THREADS < npairs) {
                double s = acc[i];
                for (int r = 0; r < nr; ++r) s += xs[r * d + pa[i]] * xs[r * d + pb[i]];
                acc[i] = s;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < COV_PAIRS_PER_THREAD; ++i) {
        const int p = threadIdx.x + i * COV_THREADS;
        if (p < npairs) partial[(long long)blockIdx.x * npairs + p] = acc[i];
    }
}

// C = sum partial / (n-1); Cholesky C = L L^T; W = L^{-T}; logJ = sum log L_ii.
__global__ void __launch_bounds__(COV_THREADS) factor_kernel(const double* partial, int nparts, int d, long long n,
                                                              double* W, double* logj, int* status) {
    __shared__ double A[MAX_DIM * MAX_DIM];
    __shared__ int bad;
    const int npairs = d * (d + 1) / 2;
    if (threadIdx.x == 0) bad = 0;
    for (int p = threadIdx.x; p < npairs; p += COV_THREADS) {
        double s = 0.0;
        for (int g = 0; g < nparts; ++g) s += partial[(long long)g * npairs + p];
        int ia = 0;
        while ((ia + 1) * (ia + 2) / 2 <= p) ++ia;
        const int ib = p - ia * (ia + 1) / 2;
        s /= (double)(n - 1);
        A[ia * d + ib] = s;
        A[ib * d + ia] = s;
    }
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    // right-looking Cholesky on the lower triangle, one warp
    for (int j = 0; j < d; ++j) {
        const double ajj = A[j * d + j];
        if (!(ajj > 0.0) || !isfinite(ajj)) {
            if (lane == 0) bad = 1;
            break;
        }
        const double ljj = sqrt(ajj);
        __syncwarp();
        for (int i = j + lane; i < d; i += 32) A[i * d + j] = (i == j) ? ljj : A[i * d + j] / ljj;
        __syncwarp();
        for (int i = j + 1 + lane; i < d; i += 32) {
            const double lij = A[i * d + j];
            for (int k = j + 1; k <= i; ++k) A[i * d + k] -= lij * A[k * d + j];
        }
        __syncwarp();
    }
    __syncwarp();
    if (bad) {
        if (lane == 0) { *status = 1; *logj = 0.0; }
        return;
    }
    // column c of L^{-1} by forward substitution; stored transposed: W[c][i] = Linv[i][c]
    for (int c = lane; c < d; c += 32) {
        for (int i = 0; i < d; ++i) {
            double x = 0.0;
            if (i >= c) {
                double s = (i == c) ? 1.0 : 0.0;
                for (int k = c; k < i; ++k) s -= A[i * d + k] * W[c * d + k];
                x = s / A[i * d + i];
            }
            W[c * d + i] = x;
        }
    }
    if (lane == 0) {
        double lj = 0.0;
        for (int j = 0; j < d; ++j) lj += log(A[j * d + j]);
        *logj = lj;
        *status = 0;
    }
}

__global__ void __launch_bounds__(128) whiten_kernel(Segment x, int d, const double* __restrict__ W,
                                                      const double* __restrict__ mu, double* __restrict__ y64,
                                                      float* __restrict__ tiles, unsigned int* rmax_bits,
                                                      long long nrows_padded) {
    extern __shared__ double sw[];  // d*d matrix, then d means
    double* smu = sw + d * d;
    for (int e = threadIdx.x; e < d * d; e += blockDim.x) sw[e] = W[e];
    for (int e = threadIdx.x; e < d; e += blockDim.x) smu[e] = mu ? mu[e] : 0.0;
    __syncthreads();
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int TJ = tile_refs(d);
    float rn = 0.f;
    if (row < nrows_padded) {
        const bool valid = row < x.n;
        const long long xrow = valid ? row : 0;
        float* tb = nullptr;
        int j = 0;
        if (tiles) {
            const long long tile = row / TJ;
            j = (int)(row - tile * TJ);
            tb = tiles + tile * (long long)((d + 1) * TJ);
        }
        // The centred row is read once (d strided loads per thread) into a per-thread array -- local memory, which is
        // interleaved per thread and therefore coalesced -- instead of d times d scattered loads; same arithmetic.
        double xr[MAX_DIM];
        if (valid)
            for (int a = 0; a < d; ++a) xr[a] = x.at(xrow, a) - smu[a];
        double n2 = 0.0;
        for (int c = 0; c < d; ++c) {
            double y = 0.0;
            if (valid) {
                for (int a = 0; a < d; ++a) y += xr[a] * sw[a * d + c];
                y64[row * d + c] = y;
            }
            if (tb) tb[(c + 1) * TJ + j] = (float)y;
            n2 += y * y;
        }
        if (tb) tb[j] = valid ? (float)(-0.5 * n2) : -INFINITY;   // row 0: h_j = -|y_j|^2 / 2
        if (valid) rn = __double2float_ru(sqrt(n2));
    }
    if (rmax_bits) {
        unsigned int bits = __float_as_uint(rn);  // non-negative floats order like their bit patterns
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bits = max(bits, __shfl_xor_sync(0xffffffffu, bits, o));
        if ((threadIdx.x & 31) == 0 && bits) atomicMax(rmax_bits, bits);
    }
}

}  // namespace

int cov_grid_size(int sm_count) { return 2 * sm_count; }

void launch_covariance(const Segment& a, const Segment& b, int d, const CovWorkspace& ws, int grid, cudaStream_t st,
                       int* launches) {
    const long long n = a.n + b.n;
    const long long nchunks = (n + cov_chunk_rows(d) - 1) / cov_chunk_rows(d);
    if (grid > nchunks) grid = (int)nchunks;
    if (grid < 1) grid = 1;
    colsum_kernel<<<grid, COV_THREADS, 0, st>>>(a, b, d, ws.partial);
    mean_kernel<<<1, MAX_DIM, 0, st>>>(ws.partial, grid, d, n, ws.mean);
    scatter_kernel<<<grid, COV_THREADS, 0, st>>>(a, b, d, ws.mean, ws.partial);
    factor_kernel<<<1, COV_THREADS, 0, st>>>(ws.partial, grid, d, n, ws.W, ws.logj, ws.status);
    *launches += 4;
}

void launch_whiten(const Segment& x, int d, const double* W, const double* mu, double* y64, float* tiles,
                   unsigned int* rmax_bits, cudaStream_t st, int* launches) {
    long long nrows = x.n;
    const int TJ = tile_refs(d);
    if (tiles) nrows = (x.n + TJ - 1) / TJ * TJ;
    if (nrows == 0) return;
    const int threads = 128;
    const long long grid = (nrows + threads - 1) / threads;
    const size_t smem = (size_t)(d * d + d) * sizeof(double);
    whiten_kernel<<<(unsigned)grid, threads, smem, st>>>(x, d, W, mu, y64, tiles, rmax_bits, nrows);
    *launches += 1;
}

}  // namespace mce
