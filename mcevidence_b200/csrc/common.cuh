// Shared declarations of the kNN-evidence kernels (sm_100a).
//
// Data layout in HBM (per problem, all owned by the context workspace):
//   X      raw samples, FP64, (n, ld) row-major as handed over by the caller
//   Y64    whitened samples, FP64, (n, d) row-major (rows are gathered by the re-rank)
//   TILES  whitened reference set for the filter, FP32, one contiguous block per tile of
//          tile_refs(d) reference rows: (d+1) rows of tile_refs(d) floats, row 0 holds
//          h_j = -|y_j|^2/2, row c+1 holds coordinate c.  A tile is one TMA bulk copy.
//   CAND   candidate indices int32 and final thresholds float written by the filter, read by the
//          re-rank: plane 0 (reference chunk 0) covers every query, [list_len][nq_pad]; planes
//          1.. only the queries of the split blocks, [ref_chunks-1][list_len][ntail_pad]
//   DKNN   FP64 (nq, K) ascending neighbour distances, IDX int64 (nq, K)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mce {

// reference rows per TMA tile: 256 while a tile stays <= 11.5 KB (ndim <= 10), else 128
__host__ __device__ constexpr int tile_refs(int d) { return d <= 10 ? 256 : 128; }
constexpr int FILTER_THREADS = 128;   // threads per filter CTA (4 warps)
constexpr int FILTER_STAGES = 4;      // TMA ring depth
constexpr int KCAP = 16;              // == MCE_MAX_NEIGHBOURS
#define MCE_TC_LIST_MAX 20             // tcgen05 engine: longest candidate list (KCAP + margin <= 4)
constexpr int MAX_DIM = 64;           // == MCE_MAX_DIM

struct Segment {            // one strided sample array, float64 or float32 (MCE_FLAG_F32_SAMPLES)
    const void* x;
    long long n;
    long long ld;           // row stride in elements
    int f32;
    const long long* rows;  // optional row map (burn-in / thinning / split applied on the device): row r = x[rows[r]]
    __host__ __device__ __forceinline__ double at(long long r, int c) const {
        const long long rr = rows ? rows[r] : r;
        return f32 ? (double)static_cast<const float*>(x)[rr * ld + c] : static_cast<const double*>(x)[rr * ld + c];
    }
    __host__ __device__ __forceinline__ Segment rows_from(long long r0, long long count) const {
        Segment s = *this;
        if (rows) s.rows = rows + r0;
        else
            s.x = f32 ? static_cast<const void*>(static_cast<const float*>(x) + r0 * ld)
                      : static_cast<const void*>(static_cast<const double*>(x) + r0 * ld);
        s.n = count;
        return s;
    }
};

// ---- K0: covariance, Cholesky, whitening matrix -------------------------------------------
struct CovWorkspace {
    double* partial;        // [cov_grid][max(d, npairs)]
    double* mean;           // [d]
    double* W;              // [d*d]  y = (x - mean) W,  W = L^{-T}
    double* logj;           // [1]    0.5 log det C
    int*    status;         // [1]    0 ok, 1 not positive definite
};
int cov_grid_size(int sm_count);
void launch_covariance(const Segment& a, const Segment& b, int d, const CovWorkspace& ws, int grid,
                       cudaStream_t st, int* launches);

// ---- K1: whitening on load ------------------------------------------------------------------
// Writes Y64 rows [0, n) and, when tiles != nullptr, the FP32 tiles (padding rows of the last
// tile get coordinates 0 and h = -inf) and atomically maxes *rmax_bits with |y| rounded up.
void launch_whiten(const Segment& x, int d, const double* W, const double* mu, double* y64,
                   float* tiles, unsigned int* rmax_bits, cudaStream_t st, int* launches);

// ---- K2: tiled FP32 filter ------------------------------------------------------------------
struct FilterParams {
    const double* yq;        // whitened queries FP64 (row r at yq + r*d); rows [q_begin, q_end) are searched
    const float*  tiles;     // reference tiles
    const int* status;       // device flag of K0: non-zero (covariance not positive definite) -> return at once
    int*   cand_idx;         // chunk 0: [list_len][nq_pad]; chunk r >= 1: + cand_off(r, e) + (ql - tail_q0)
    float* cand_th;          // chunk 0: [nq_pad];           chunk r >= 1: nq_pad + (r-1) ntail_pad + (ql - tail_q0)
    long long q_begin, q_end;
    long long nq_pad;
    long long ntail_pad;     // padded number of queries in the split blocks
    long long tail_q0;       // first local query id of the split blocks (= nfull * queries per block)
    long long self_offset;   // auto mode: reference index of query row r is r + self_offset; cross mode: any
    int self_exclude;        // 1 in auto mode
    int ntiles;
    int tiles_per_chunk;     // tiles per reference chunk of a split query block
    int ref_chunks;          // chunks of a split query block
    int nqblocks;
    int nfull;               // query blocks [0, nfull) scan the whole reference set as one unit (chunk 0);
                             // blocks [nfull, nqblocks) are split into ref_chunks units each
    int list_len;
    int d;
    const int* chunk_tile0;  // batch-size sweep: ref_chunks + 1 tile indices, chunk r = tiles [chunk_tile0[r], chunk_tile0[r+1])
                             // (the prefix boundaries of MCEvidence.py:1039-1045); nullptr: uniform tiles_per_chunk
    int chunk_base;          // sweep: this launch covers reference chunk chunk_base + (unit's chunk) only -- one launch per
                             // chunk, in order, so that chunk c can start from the final thresholds of chunk c - 1
    int seed_th;             // sweep: 1 = a unit of chunk c >= 1 starts from the threshold plane of chunk c - 1 as a floor.
                             // Valid because batches are nested prefixes: whoever reads chunk c also reads chunks < c, and
                             // L references with a score >= that floor are already in their lists.
};
// first and one-past-last tile of reference chunk r of a split query block
__device__ __forceinline__ void chunk_tile_range(const FilterParams& p, int r, int* t0, int* t1) {
    if (p.chunk_tile0) {
        *t0 = p.chunk_tile0[r];
        *t1 = p.chunk_tile0[r + 1];
    } else {
        *t0 = r * p.tiles_per_chunk;
        *t1 = min(*t0 + p.tiles_per_chunk, p.ntiles);
    }
}
// element offsets into the candidate buffers (shared by the filter and the re-rank)
__host__ __device__ __forceinline__ long long cand_idx_offset(int r, int e, long long ql, int list_len, long long nq_pad,
                                                              long long ntail_pad, long long tail_q0) {
    return r == 0 ? (long long)e * nq_pad + ql
                  : (long long)list_len * nq_pad + ((long long)(r - 1) * list_len + e) * ntail_pad + (ql - tail_q0);
}
__host__ __device__ __forceinline__ long long cand_th_offset(int r, long long ql, long long nq_pad, long long ntail_pad,
                                                             long long tail_q0) {
    return r == 0 ? ql : nq_pad + (long long)(r - 1) * ntail_pad + (ql - tail_q0);
}
struct FilterConfig { int tq, tr, packed, sync; };
// Returns 0 and fills grid/smem/ctas_per_sm on success, -1 if (d, cfg) has no instantiation.
int filter_query(int d, const FilterConfig& cfg, int list_len, int* queries_per_cta, int* smem_bytes,
                 int* ctas_per_sm);
int launch_filter(const FilterParams& p, const FilterConfig& cfg, cudaStream_t st, int* launches);
void filter_default_config(int d, int list_len, FilterConfig* cfg);

// ---- K2 (tensor-core engine, opt-in): tcgen05 TF32 coarse scores + exact FP32 slow path -------
// Same unit decomposition, candidate planes and threshold semantics as the FFMA filter, so K3 is
// shared.  Reference image: tiles of TC_TILE_REFS rows, each K-major in the UMMA "interleaved"
// (no-swizzle) canonical layout [K/4 chunks][16 row groups][8 rows][4 floats]; column d and d+1 of
// a row hold h split into a TF32-exact head and the remainder, the other padding columns are 0.
constexpr int TC_TILE_REFS = 128;     // reference rows per tile (UMMA N)
constexpr int TC_QUERIES = 256;       // queries per CTA: two UMMA M=128 blocks
__host__ __device__ constexpr int tc_kpad(int d) { return (d + 2 + 7) / 8 * 8; }   // K: coordinates + (h_hi, h_lo), multiple of 8
struct TcFilterParams {
    FilterParams f;                   // f.tiles = TC image; f.ntiles / f.tiles_per_chunk count TC tiles
    const unsigned int* rmax_bits;    // max |y| over the references (written by K1)
    float slack_qr, slack_rr;         // coarse-score error bound e = slack_qr |q| Rmax + slack_rr Rmax^2
    unsigned int* debug;              // CHECK builds: [0] max over pairs of |coarse - exact| / e as float bits, [1] pairs checked (mod 2^32)
    long long* timeline;              // timing diagnostics: clock64() of the hand-over points of CTA 0's first tiles, or null
    float* dump;                      // CHECK builds: coarse scores of unit 0, tile 0: [TC_QUERIES][TC_TILE_REFS] (may be null)
    int stages;                       // shared-memory ring depth
    int swap_offsets;                 // debug bits: 1 = exchange the leading/stride byte offsets of the descriptors, 2 = version field 0
    int check;                        // 1: evaluate every pair exactly as well and record the worst ratio in debug[]
    int mma_warps;                    // 1 or 2 warps issuing the MMAs (2: even / odd tiles)
    int perf_mode;                    // timing experiments only (results are garbage): 1 = epilogue skips the TMEM loads
                                      // (MMA + TMA pipeline alone), 2 = loads without the threshold test
};
size_t tc_image_bytes(long long n_ref, int d);
// Builds the TC image from the whitened FP64 rows and the h row of the FFMA tiles (bitwise the same h).
void launch_tc_image(const double* y64, const float* ffma_tiles, long long n_ref, int d, float* image, cudaStream_t st,
                     int* launches);
// Returns 0 and fills the launch shape on success, -1 if the configuration cannot run.
// want_stages <= 0: as many ring slots (2..8) as fit the shared-memory budget
int tc_filter_query(int d, int list_len, int want_stages, int* queries_per_cta, int* smem_bytes, int* ctas_per_sm, int* stages);
int launch_tc_filter(const TcFilterParams& p, cudaStream_t st, int* launches);
// {flag, a, b, block} of a barrier wait that gave up (0 flag = none); clears the record
int tc_read_stuck(unsigned int (&out)[4], cudaStream_t st);

// ---- K3: FP64 re-rank + verification, exact re-scan ---------------------------------------
struct RerankParams {
    const double* yq;        // whitened queries
    const double* yr;        // whitened references
    const int*   status;     // as in FilterParams: skip all work after a failed factorisation
    const int*   cand_idx;
    const float* cand_th;
    long long ntail_pad, tail_q0;   // candidate planes of chunks >= 1 (see FilterParams)
    const unsigned int* rmax_bits;
    double*    dknn;         // (nq, out_ld) ; neighbours written at column out_col0..
    long long* idx;
    int* flagged;            // [nq] local row ids needing an exact re-scan
    int* nflagged;
    long long q_begin, q_end, nq_pad;
    long long n_ref;
    long long self_offset;
    int self_exclude;
    int ref_chunks, list_len;
    int nfull_queries;       // queries [0, nfull_queries) have one candidate chunk, the others ref_chunks
    int d, k;                // k true neighbours wanted
    int out_ld, out_col0;    // output row stride and first column (1 when a synthetic self column is kept)
    int write_self_col;      // write (0.0, own index) into column 0
    int flag_all;            // MCE_FLAG_EXACT_ONLY: flag every row
    // Batch-size sweep (MCEvidence.py:1039-1045; sweep_nb == 0: one problem).  Batch b searches queries [0, S_b) among
    // references [0, S_b) (auto mode).  One filter pass over the largest batch with reference chunks cut at the tile
    // aligned batch sizes serves every batch: batch b merges the candidates of the chunks wholly inside [0, S_b) and
    // scans the references of the partial tile [sweep_edge[b], S_b) directly.  Output row of (b, query j) =
    // sweep_off[b] + j; that row id is also what `flagged` holds.
    const long long* sweep_off;    // [nb + 1]
    const long long* sweep_rows;   // [nb]  S_b
    const long long* sweep_edge;   // [nb]  first reference not covered by the chunks of batch b (tile aligned, <= S_b)
    const int* sweep_nchunk;       // [nb]  chunks wholly inside [0, S_b)
    int sweep_nb;
    long long sweep_total;         // = sweep_off[nb] (host copy: the launch size)
};
// sweep row id -> (batch, query)
__device__ __forceinline__ int sweep_batch_of(const RerankParams& p, long long t) {
    int b = 0;
    while (b + 1 < p.sweep_nb && t >= p.sweep_off[b + 1]) ++b;
    return b;
}
void launch_rerank(const RerankParams& p, cudaStream_t st, int* launches);
size_t exact_scratch_bytes(int grid);
void launch_exact_rescan(const RerankParams& p, int grid, void* scratch, cudaStream_t st, int* launches);

// ---- K4: log-volume + logsumexp epilogue ---------------------------------------------------
struct ReduceParams {
    const double* dknn;      // (nq, ld) neighbour distances; neighbour i (1-based) at column col0 + i - 1
    const double* logl;      // indexed by global query row
    const double* weights;
    double* block_partial;   // [grid][nk][3]
    double* out;             // [nk][3]  (max, sum, sum2)
    long long q_begin, q_end;
    int ld, col0;
    int first_nb;            // neighbour used by output 0: 1 (auto) or 2 (cross)
    int nk;                  // kmax - 1 outputs
    int d;
};
int reduce_grid_size(long long nq);
void launch_reduce(const ReduceParams& p, int grid, cudaStream_t st, int* launches);

// ---- FP32 FMA microbenchmark ----------------------------------------------------------------
void launch_fma_peak(int variant, int iters, int grid, float* sink, cudaStream_t st);
long long fma_peak_flops(int variant, int iters, int grid);

// ---- unit peaks of the tensor-core engine (tc_peak.cu): 0 FMNMX3, 1/2 tcgen05.ld (8/4 warps), 3 tcgen05.mma tf32 ----
int launch_unit_peak(int unit, int iters, int sm_count, float* sink, cudaStream_t st, int* grid_out);
double unit_peak_work(int unit, int iters, int grid);

}  // namespace mce
