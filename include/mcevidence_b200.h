/*
 * mcevidence_b200 -- C ABI of the B200 kNN-evidence hot path.
 *
 * The reference (yabebalFantaye/MCEvidence) has no FFI: its numeric core is the
 * body of MCEvidence.evidence() (MCEvidence.py:1033-1131) plus scikit-learn's
 * NearestNeighbors.  This header declares the entry points a maintainer binds
 * with ctypes in place of those lines:
 *
 *   mce_evidence   replaces  get_covariance + diagonalise_chain + NearestNeighbors.fit /
 *                            kneighbors + the volume/dotp/MLE loop
 *                            (MCEvidence.py:851-882, 842-849, 1093-1104, 1107-1131)
 *   mce_knn        replaces  NearestNeighbors(...).fit(ref).kneighbors(query)
 *                            (MCEvidence.py:1093-1104) with the same (DkNN, indices) contract
 *   mce_evidence_batch       the per-(model, data, chain) loop of planck_mcevidence.py:306-359
 *
 * Plain C types only; every pointer is caller-owned and only has to stay valid
 * for the duration of the call.  Functions return MCE_OK (0) or a negative code
 * and never throw or exit; mce_last_error() gives the message (thread-local).
 * One call at a time per context (internal mutex); contexts are independent.
 * There is no CPU fallback: without a usable sm_100 device every compute entry
 * point fails with MCE_ERR_CUDA.
 */
#ifndef MCEVIDENCE_B200_H
#define MCEVIDENCE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCE_ABI_VERSION 2

#define MCE_OK              0
#define MCE_ERR_ARG        -1   /* bad argument (message says which) */
#define MCE_ERR_CUDA       -2   /* CUDA runtime / driver error, or no device */
#define MCE_ERR_NOT_POSDEF -3   /* sample covariance is not positive definite (MCEvidence.py:867) */
#define MCE_ERR_NOMEM      -4   /* device or host allocation failed */
#define MCE_ERR_INTERNAL   -5   /* exactness verification failed after the exact re-scan */

#define MCE_MAX_DIM        64   /* largest ndim served by the tiled FP32 filter */
#define MCE_MAX_NEIGHBOURS 16   /* largest number of true neighbours per sample */

/* mce_problem.flags */
#define MCE_FLAG_DEVICE_PTRS  1  /* query/ref/logl/weights and result arrays are device pointers on the context's device */
#define MCE_FLAG_PARTIALS     2  /* also return the per-k (max, sum, sum2) partials (result.partials) */
#define MCE_FLAG_EXACT_ONLY   4  /* skip the FP32 filter: exhaustive FP64 search for every row (validation) */
#define MCE_FLAG_F32_SAMPLES  8  /* query and ref point to float32 arrays (strides in elements); every other array
                                    stays float64.  The values are widened on load: results equal those of the same
                                    call on the widened float64 copy, bit for bit */

typedef struct mce_ctx mce_ctx;

/* One evidence problem = one batch of MCEvidence.evidence() (MCEvidence.py:1041-1131). */
typedef struct mce_problem {
    const void   *query;     /* s1 samples, float64 (or float32 with MCE_FLAG_F32_SAMPLES), (n_query, ld_query) row-major,
                                first d columns used (MCEvidence.py:902) */
    const void   *ref;       /* s2 samples for cross mode (MCEvidence.py:1075-1094), or NULL: auto mode,
                                neighbours searched in the query set itself with each row excluded (k0=1, :1099) */
    const double *logl;      /* n_query log-likelihoods: -loglikes, negated again for pos_lnp (MCEvidence.py:399,1060) */
    const double *weights;   /* n_query sample weights (MCEvidence.py:400,1117) */
    int64_t n_query;
    int64_t n_ref;           /* ignored when ref == NULL */
    int64_t ld_query;        /* row stride of query in elements (>= d) */
    int64_t ld_ref;
    int64_t q_begin;         /* only rows [q_begin, q_end) of query are searched and reduced (row sharding over */
    int64_t q_end;           /* GPUs); covariance and, in auto mode, the reference set still use all rows. 0,0 = all */
    int32_t d;               /* ndim */
    int32_t kmax;            /* reference kmax (clamped to >= 2, MCEvidence.py:694): logE for k = 1..kmax-1 */
    int32_t flags;
    int32_t reserved;
    /* Whitening y = (x - mu) W.  W == NULL: the library estimates the unweighted, ddof=1 covariance of
       query U ref (np.cov, MCEvidence.py:865), factors it (Cholesky) and sets log J = 0.5 log det C (:877).
       Otherwise W is d x d row-major (for the reference's own factor: eVec / sqrt(eVal), :844-847),
       W_ref optionally a different matrix for the ref set (covtype='single', :1080-1086), mu may be NULL (0).
       W, W_ref and mu are always HOST pointers, also with MCE_FLAG_DEVICE_PTRS. */
    const double *W;
    const double *W_ref;
    const double *mu;
    double log_jacobian;     /* log J, read only when W != NULL */
    double log_sum_w;        /* log sum of adjusted_weights of s1 (MCEvidence.py:1126) */
    double log_prior_volume; /* MCEvidence.py:1017-1020 */
    /* Burn-in / thinning / split applied on the device (MCEvidence.py:174-249, 350-391, 417-445): when query_rows is
       not NULL, `query` points to a BASE array of n_query_base rows (the concatenated chain) and sample i of s1 is its
       row query_rows[i], i < n_query -- the rows kept by burn-in, thinning and the random split, in the order the
       reference's fancy indexing would produce.  The base array is uploaded once and the rows are gathered by the
       covariance and whitening kernels as they load; no gathered copy is made on either side.  ref_rows likewise for
       `ref`; when ref == query (same base pointer and stride: the two halves of one split chain) the base is
       uploaded once.  logl and weights stay per-sample host vectors of n_query entries (thinning replaces the weights,
       :427-445).  NULL, NULL, 0, 0: the arrays are the sample sets themselves.  Host-array problems only. */
    const int64_t *query_rows;
    const int64_t *ref_rows;
    int64_t n_query_base;
    int64_t n_ref_base;
} mce_problem;

typedef struct mce_stats {
    double ms_total;         /* device time of the whole call (CUDA events on the context stream) */
    double ms_h2d;           /* host->device copies (0 with MCE_FLAG_DEVICE_PTRS) */
    double ms_cov;           /* K0 covariance + Cholesky */
    double ms_whiten;        /* K1 whitening + FP32 tile emission */
    double ms_filter;        /* K2 tiled FP32 filter (the dominant kernel) */
    double ms_rerank;        /* K3 FP64 re-rank + verification */
    double ms_exact;         /* exact FP64 re-scan of unverified rows */
    double ms_reduce;        /* K4 log-volume + logsumexp epilogue */
    double ms_d2h;
    int64_t pairs;           /* pair distances evaluated by the filter = (q_end-q_begin) * n_ref */
    int64_t rows_rescanned;  /* rows that failed the FP32->FP64 verification and were re-scanned exactly */
    int32_t launches;        /* kernels launched by this call */
    int32_t filter_tq;       /* filter configuration actually used */
    int32_t filter_tr;
    int32_t filter_packed;
    int32_t ref_chunks;
    int32_t list_len;
    int32_t filter_grid;
    int32_t filter_smem;
    int32_t filter_ctas_per_sm;
    float   rmax;            /* largest whitened reference norm (enters the error bound) */
    int32_t filter_mode;     /* 0 free-running warps, 1 CTA barrier per tile group, 2+ with register prefetch; 16 = tcgen05 engine */
} mce_stats;

typedef struct mce_result {
    double  *logE;           /* kmax-1 values: MLE[0,1:] of MCEvidence.py:1155-1157 (may be NULL with MCE_FLAG_PARTIALS) */
    double  *var_logE;       /* kmax-1 delta-method variances of logE, or NULL (no reference counterpart) */
    double  *dknn;           /* optional (q_end-q_begin, kmax+1) distances shaped like kneighbors() at MCEvidence.py:1104 */
    int64_t *idx;            /* optional neighbour indices, same shape */
    double  *partials;       /* with MCE_FLAG_PARTIALS: 3*(kmax-1) doubles (max_t, sum exp(t-max), sum exp(2(t-max))) */
    double   log_jacobian;   /* log J actually used */
    mce_stats stats;
} mce_result;

int         mce_version(void);
const char *mce_last_error(void);
int         mce_device_count(void);

int  mce_ctx_create(int device, mce_ctx **out);
void mce_ctx_destroy(mce_ctx *ctx);
/* Run on an existing CUDA stream (e.g. torch's current stream) instead of the context's own non-blocking
   stream; pass cudaStreamLegacy ((void*)1) for the default stream.  NULL restores the context's own stream. */
int  mce_ctx_set_stream(mce_ctx *ctx, void *cuda_stream);
/* Tuning knobs: "filter_tq", "filter_tr", "filter_packed", "filter_sync" (kernel mode 0/1/2), "ref_chunks"
   (uniform split of the reference set), "list_margin". -1 = automatic.  "batch_lanes" (1..32, default 8): concurrent
   problems of mce_evidence_batch.  "batch_stage" (default 1): mce_evidence_batch packs the used columns of the host arrays
   into a pinned arena on a few host threads so that the uploads overlap the kernels; 0 = copy from the caller's
   (pageable) arrays directly.  "tc_min_mpairs": the automatic engine choice takes the tcgen05 filter from
   this many million pairs on (0 = the measured default: 200 for ndim <= 6, 300 above).  "sweep_seed" (default 1): mce_evidence_sweep launches the filter once per batch
   boundary and seeds every range with the thresholds of the one before; 0 = one launch with restarted lists.
   "filter_engine": 2 = chosen per problem (default: the tcgen05 filter for ndim >= 4 and at least ~3e8 pairs, 2e8
   up to ndim 6 -- where it was measured faster on B200 -- otherwise the FP32 FFMA filter), 0 = FFMA filter,
   1 = tcgen05 tensor-core filter (TF32 coarse scores in TMEM, exact FP32 evaluation of the columns that pass; same
   candidates contract, stats.filter_mode reports 16); the environment variable MCE_FILTER_ENGINE=0/1/2 sets the
   initial value of every new context.
   Diagnostics of the tcgen05 engine: "tc_check" (1: also evaluate every pair exactly and record the worst
   |coarse - exact| / allowance, read back with mce_ctx_get_debug), "tc_slack_percent" (scales the allowance,
   default 100), "tc_swap_offsets" (bit 0: exchange the two byte offsets of the shared-memory matrix descriptors,
   bit 1: descriptor version field 0), "tc_stages" (shared-memory ring depth 2..8, -1 = as many as fit),
   "tc_mma_warps" (1 or 2 warps issuing the MMAs, default 2), "tc_timeline" (ndim 7..14: record the hand-over
   times of CTA 0's first 1024 tiles, read back as "tc_dump"), "tc_perf_mode" (ndim 7..14, timing experiments whose
   results are garbage: 1 = the epilogue skips the TMEM loads, 2 = loads without the threshold test, 3 = loads and
   threshold test but nothing ever passes, i.e. the scan without a single hit). */
int  mce_ctx_set_option(mce_ctx *ctx, const char *name, int value);
/* Diagnostics of the last call on this context: "tc_check_ratio" (worst |coarse - exact| / allowance seen by a
   "tc_check" run; must stay below 1), "tc_check_pairs" (pairs compared, modulo 2^32), "filter_engine". */
int  mce_ctx_get_debug(mce_ctx *ctx, const char *name, double *value);
/* "tc_dump": after a "tc_check" run, the coarse scores of the first unit's first tile, float[256 queries][128
   references]; *n receives the element count, `out` may be NULL to query it. */
int  mce_ctx_get_debug_array(mce_ctx *ctx, const char *name, float *out, int64_t capacity, int64_t *n);

int  mce_evidence(mce_ctx *ctx, const mce_problem *problem, mce_result *result);
/* n independent problems (planck_mcevidence.py:306-359).  Host-array problems overlap on up to "batch_lanes" (option,
   default 8) internal streams with their own workspaces: inputs are uploaded and kernels enqueued for every problem
   without waiting for results, one synchronisation at the end; results are bit-identical to n mce_evidence calls.
   stats of a batched problem carry only ms_total and the configuration fields. */
int  mce_evidence_batch(mce_ctx *ctx, int n, const mce_problem *problems, mce_result *results);
/* Batch-size sweep (set_batch / the batch loop of MCEvidence.py:808-840, 1039-1045): results[b] = the evidence of the
   first batch_rows[b] rows of `query` (auto mode: neighbours among those rows; cross mode: in all of ref), with the
   whitening of the whole problem (W == NULL: covariance of ALL n_query rows U ref, as covtype='all' does), SumW =
   log_sum_w for every batch and S = batch_rows[b] in the normalisation.  One upload, one whitening and one pass of the
   filter over the largest batch serve every size: the reference set is cut at the tile-aligned batch sizes, batch b
   merges the candidates of the ranges inside its prefix in the FP64 re-rank.  Each result equals the mce_evidence call
   on the prefix with the same explicit W / mu, bit for bit.  q_begin/q_end must be 0; dknn/idx are not returned. */
int  mce_evidence_sweep(mce_ctx *ctx, const mce_problem *problem, int nbatch, const int64_t *batch_rows, mce_result *results);

/* Device group: the GPUs of one box driven from ONE process, one host thread per device (the multi-GPU form of the
   entry points above).  device_ids == NULL: devices 0..n_devices-1; n_devices <= 0: all.
   mce_group_evidence shards the query rows over the devices: the sample arrays are uploaded once to the first device
   and sent to the others with ncclBroadcast over NVLink (NCCL is bound with dlopen("libnccl.so.2"); when it cannot be
   loaded, or MCE_GROUP_NO_NCCL=1, every device uploads the host arrays itself), every device whitens the whole
   reference set and searches its block of rows, and the 3 (kmax-1) partial sums per device -- already in this
   process's host memory -- are combined with mce_combine_partials.  mce_group_evidence_batch deals independent
   problems largest-first over the devices (chain-per-GPU, planck_mcevidence.py:149-160) and runs each share as one
   mce_evidence_batch.  Host-array problems only. */
typedef struct mce_group mce_group;
int  mce_group_create(int n_devices, const int *device_ids, mce_group **out);
void mce_group_destroy(mce_group *group);
int  mce_group_size(const mce_group *group);
int  mce_group_uses_nccl(const mce_group *group);
int  mce_group_set_option(mce_group *group, const char *name, int value);
int  mce_group_evidence(mce_group *group, const mce_problem *problem, mce_result *result);
int  mce_group_evidence_batch(mce_group *group, int n, const mce_problem *problems, mce_result *results);

/* k nearest neighbours with sklearn's contract: dknn/idx are (q_end-q_begin, n_neighbors), ascending.
   Auto mode (ref == NULL): column 0 is the row itself at distance 0, as when the reference queries the
   fitted set with itself (MCEvidence.py:1101-1104).  Only the sample arrays, their sizes and strides, the row range, d, the whitening fields and
   flags are read. */
int  mce_knn(mce_ctx *ctx, const mce_problem *problem, int n_neighbors, double *dknn, int64_t *idx, mce_stats *stats);

/* Host arithmetic shared by the single- and multi-GPU paths.
   Combine `nparts` partial triples (each 3*nk doubles) into one. */
int  mce_combine_partials(int nparts, int nk, const double *parts, double *out);
/* logE_k = log SumW + log J - log PV - log(S k_nn + 1) + log c_d + max_k + log sum_k  (MCEvidence.py:1117-1131) */
int  mce_finalize(const double *partials, int nk, int d, int cross, int64_t n_query_total, double log_sum_w,
                  double log_jacobian, double log_prior_volume, double *logE, double *var_logE);

/* The work-decomposition model of the filter (host arithmetic, no device needed; for inspection and tests):
   how many query blocks run as whole-set units (`nfull`) and into how many reference ranges the remaining
   blocks are split (`chunks`) for a GPU with `sm_count` SMs holding `ctas_per_sm` CTAs of `queries_per_cta`
   queries; `wide` selects the constants of the four-queries-per-thread tile.  `cost` (optional) is the
   modelled time in comparable units. */
int  mce_plan_split(int sm_count, int d, int64_t n_query, int64_t n_ref, int list_len, int queries_per_cta,
                    int ctas_per_sm, int wide, int *nfull, int *chunks, double *cost);

/* FP32 FMA microbenchmark on the context's device: variant 0 = scalar FFMA, 1 = packed FFMA2 (2-4: other
   FFMA2 operand-sharing patterns, see csrc/fma_peak.cu).
   Reports achieved TFLOP/s counting 2 FLOP per lane-FMA, the roofline denominator of the filter. */
int  mce_fma_peak(mce_ctx *ctx, int variant, int iters, double *tflops, double *ms);

/* Unit peaks behind the tcgen05 engine, measured with the instruction shapes it uses: unit 0 = FMNMX3
   (lane-operations/s), 1 / 2 = tcgen05.ld 32x32b.x32 with 8 / 4 warps per SM (bytes/s), 3 = tcgen05.mma
   kind::tf32 M=128 N=128 K=8 from shared memory (FLOP/s). */
int  mce_unit_peak(mce_ctx *ctx, int unit, int iters, double *rate, double *ms);

/* Host-side chain ingestion (next-row N1): parse a whitespace-separated text chain (`#` comments, blank
   lines skipped) with `nthreads` threads (<= 0: all cores) into a malloc'ed row-major (nrows, ncols) array --
   the np.loadtxt call of MCEvidence.py:564, same correctly-rounded values.  Free with mce_free_host. */
int  mce_read_chain_text(const char *path, int nthreads, double **data, int64_t *nrows, int64_t *ncols);
void mce_free_host(void *p);

#ifdef __cplusplus
}
#endif
#endif /* MCEVIDENCE_B200_H */
