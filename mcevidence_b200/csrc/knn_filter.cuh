// K2: tiled brute-force FP32 filter (the dominant kernel).
//
// Replaces the neighbour search of NearestNeighbors.kneighbors (MCEvidence.py:1104; sklearn
// kd_tree / ArgKmin) by an exhaustive scan whose survivors are re-ranked exactly in FP64 (K3).
//
// Work decomposition: one CTA = one (query block, reference chunk) unit.  A query block is
// FILTER_THREADS * TQ whitened queries; each thread keeps TQ queries (TQ*D floats) in
// registers for the whole unit.  The reference chunk is streamed as tiles of TILE_REFS rows
// through a FILTER_STAGES-deep shared-memory ring filled by TMA bulk copies
// (cp.async.bulk + mbarrier complete_tx; one contiguous (D+1)*TILE_REFS*4-byte block per tile).
//
// Ranking score: s(q, r) = q.r - |r|^2/2 = (|q|^2 - |q - r|^2) / 2, larger is closer.  The
// accumulator is initialised with h_r = -|r|^2/2 (an extra tile row), so one pair costs exactly
// D FMAs.  In the packed variant two references share one FFMA2 (fma.rn.f32x2): the query
// coordinate is a scalar register broadcast to both halves by the instruction, the reference
// pair comes straight out of a broadcast LDS.128.
//
// Selection: each thread compares the TR scores of a step against its queries' thresholds
// (one FSETP per pair, or-reduced to a single rarely taken branch).  Passing scores are
// inserted into a per-(thread, query) list of list_len entries in shared memory (replace the
// minimum, new threshold = new minimum).  A thread sees every reference of the chunk for its
// own queries, so no cross-thread merge is needed; chunks are merged by the re-rank.
#pragma once
#include "common.cuh"
#include <math_constants.h>

namespace mce {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}

// Insert (s, idx) into the list of one (thread, query): entries are strided by FILTER_THREADS.
// The caller guarantees s > current minimum.  Returns the new minimum (the new threshold).
static __device__ __noinline__ float list_insert(float* ls, int* li, int len, float s, int idx) {
    float m = ls[0];
    int am = 0;
    float second = CUDART_INF_F;
    for (int e = 1; e < len; ++e) {
        const float v = ls[e * FILTER_THREADS];
        if (v < m) {
            second = m;
            m = v;
            am = e;
        } else if (v < second) {
            second = v;
        }
    }
    ls[am * FILTER_THREADS] = s;
    li[am * FILTER_THREADS] = idx;
    return fminf(s, second);
}

template <int D, int TQ, int TR, bool PACKED>
__global__ void __launch_bounds__(FILTER_THREADS) knn_filter_kernel(const FilterParams p) {
    constexpr int NT = FILTER_THREADS;
    constexpr int TJ = TILE_REFS;
    constexpr int NS = FILTER_STAGES;
    constexpr int TILE_FLOATS = (D + 1) * TJ;
    constexpr uint32_t TILE_BYTES = TILE_FLOATS * sizeof(float);
    static_assert(TR == 4 || TR == 8, "TR must be 4 or 8");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stage = reinterpret_cast<float*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NS * TILE_BYTES);
    float* ls = reinterpret_cast<float*>(bars + NS);
    const int L = p.list_len;
    int* li = reinterpret_cast<int*>(ls + TQ * L * NT);

    const int tid = threadIdx.x;
    const int qblock = blockIdx.x % p.nqblocks;
    const int rchunk = blockIdx.x / p.nqblocks;
    const long long ql0 = (long long)qblock * (NT * TQ) + tid;  // local id of this thread's first query
    const long long nq = p.q_end - p.q_begin;
    const int t0 = rchunk * p.tiles_per_chunk;
    const int t1 = min(t0 + p.tiles_per_chunk, p.ntiles);

    // queries -> registers (FP64 whitened rows rounded to FP32), thresholds, empty lists
    float q[TQ][D];
    float th[TQ];
#pragma unroll
    for (int t = 0; t < TQ; ++t) {
        const long long ql = ql0 + (long long)t * NT;
        const bool valid = ql < nq;
        const double* src = p.yq + (p.q_begin + (valid ? ql : 0)) * D;
#pragma unroll
        for (int c = 0; c < D; ++c) q[t][c] = valid ? (float)src[c] : 0.f;
        th[t] = valid ? -CUDART_INF_F : CUDART_INF_F;
        for (int e = 0; e < L; ++e) {
            ls[(t * L + e) * NT + tid] = -CUDART_INF_F;
            li[(t * L + e) * NT + tid] = -1;
        }
    }
    const int self0 = p.self_exclude ? (int)(p.q_begin + ql0 + p.self_offset) : -0x7fffffff;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) mbar_init(smem_u32(&bars[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const float* gt = p.tiles + (long long)t0 * TILE_FLOATS;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NS - 1; ++s) {
            if (t0 + s < t1) {
                const uint32_t bar = smem_u32(&bars[s]);
                mbar_arrive_expect_tx(bar, TILE_BYTES);
                tma_bulk_g2s(smem_u32(stage + s * TILE_FLOATS), gt + (long long)s * TILE_FLOATS, TILE_BYTES, bar);
            }
        }
    }

    for (int it = 0; it < t1 - t0; ++it) {
        const int slot = it % NS;
        const uint32_t parity = (it / NS) & 1;
        if (tid == 0) {
            const int nxt = it + NS - 1;
            if (t0 + nxt < t1) {
                const int ns = nxt % NS;  // the slot released by the __syncthreads of iteration it-1
                const uint32_t bar = smem_u32(&bars[ns]);
                mbar_arrive_expect_tx(bar, TILE_BYTES);
                tma_bulk_g2s(smem_u32(stage + ns * TILE_FLOATS), gt + (long long)nxt * TILE_FLOATS, TILE_BYTES, bar);
            }
        }
        {
            const uint32_t bar = smem_u32(&bars[slot]);
            while (!mbar_try_wait(bar, parity)) {
            }
        }
        const float* tile = stage + slot * TILE_FLOATS;
        const int jbase = (t0 + it) * TJ;

#pragma unroll 1
        for (int j0 = 0; j0 < TJ; j0 += TR) {
            bool any = false;
            if constexpr (PACKED) {
                float2 acc[TQ][TR / 2];
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    float2 rv[TR / 2];
#pragma unroll
                    for (int g = 0; g < TR / 4; ++g) {
                        const float4 r4 = *reinterpret_cast<const float4*>(tile + c * TJ + j0 + 4 * g);
                        rv[2 * g] = make_float2(r4.x, r4.y);
                        rv[2 * g + 1] = make_float2(r4.z, r4.w);
                    }
                    if (c == 0) {
                        float2 hv[TR / 2];
#pragma unroll
                        for (int g = 0; g < TR / 4; ++g) {
                            const float4 h4 = *reinterpret_cast<const float4*>(tile + D * TJ + j0 + 4 * g);
                            hv[2 * g] = make_float2(h4.x, h4.y);
                            hv[2 * g + 1] = make_float2(h4.z, h4.w);
                        }
#pragma unroll
                        for (int t = 0; t < TQ; ++t)
#pragma unroll
                            for (int pp = 0; pp < TR / 2; ++pp)
                                acc[t][pp] = __ffma2_rn(make_float2(q[t][0], q[t][0]), rv[pp], hv[pp]);
                    } else {
#pragma unroll
                        for (int t = 0; t < TQ; ++t)
#pragma unroll
                            for (int pp = 0; pp < TR / 2; ++pp)
                                acc[t][pp] = __ffma2_rn(make_float2(q[t][c], q[t][c]), rv[pp], acc[t][pp]);
                    }
                }
#pragma unroll
                for (int t = 0; t < TQ; ++t)
#pragma unroll
                    for (int pp = 0; pp < TR / 2; ++pp) any |= (acc[t][pp].x > th[t]) | (acc[t][pp].y > th[t]);
                if (any) {
#pragma unroll
                    for (int t = 0; t < TQ; ++t) {
#pragma unroll
                        for (int e = 0; e < TR; ++e) {
                            const float s = (e & 1) ? acc[t][e / 2].y : acc[t][e / 2].x;
                            if (s > th[t]) {
                                const int ridx = jbase + j0 + e;
                                if (ridx != self0 + t * NT)
                                    th[t] = list_insert(ls + t * L * NT + tid, li + t * L * NT + tid, L, s, ridx);
                            }
                        }
                    }
                }
            } else {
                float acc[TQ][TR];
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    float rv[TR];
#pragma unroll
                    for (int g = 0; g < TR / 4; ++g) {
                        const float4 r4 = *reinterpret_cast<const float4*>(tile + c * TJ + j0 + 4 * g);
                        rv[4 * g] = r4.x;
                        rv[4 * g + 1] = r4.y;
                        rv[4 * g + 2] = r4.z;
                        rv[4 * g + 3] = r4.w;
                    }
                    if (c == 0) {
                        float hv[TR];
#pragma unroll
                        for (int g = 0; g < TR / 4; ++g) {
                            const float4 h4 = *reinterpret_cast<const float4*>(tile + D * TJ + j0 + 4 * g);
                            hv[4 * g] = h4.x;
                            hv[4 * g + 1] = h4.y;
                            hv[4 * g + 2] = h4.z;
                            hv[4 * g + 3] = h4.w;
                        }
#pragma unroll
                        for (int t = 0; t < TQ; ++t)
#pragma unroll
                            for (int e = 0; e < TR; ++e) acc[t][e] = fmaf(q[t][0], rv[e], hv[e]);
                    } else {
#pragma unroll
                        for (int t = 0; t < TQ; ++t)
#pragma unroll
                            for (int e = 0; e < TR; ++e) acc[t][e] = fmaf(q[t][c], rv[e], acc[t][e]);
                    }
                }
#pragma unroll
                for (int t = 0; t < TQ; ++t)
#pragma unroll
                    for (int e = 0; e < TR; ++e) any |= (acc[t][e] > th[t]);
                if (any) {
#pragma unroll
                    for (int t = 0; t < TQ; ++t) {
#pragma unroll
                        for (int e = 0; e < TR; ++e) {
                            if (acc[t][e] > th[t]) {
                                const int ridx = jbase + j0 + e;
                                if (ridx != self0 + t * NT)
                                    th[t] = list_insert(ls + t * L * NT + tid, li + t * L * NT + tid, L, acc[t][e], ridx);
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();  // every warp is done with this slot before thread 0 refills it
    }

    // write the lists and final thresholds (coalesced over tid)
#pragma unroll
    for (int t = 0; t < TQ; ++t) {
        const long long ql = ql0 + (long long)t * NT;
        if (ql < nq) {
            p.cand_th[(long long)rchunk * p.nq_pad + ql] = th[t];
            for (int e = 0; e < L; ++e)
                p.cand_idx[((long long)rchunk * L + e) * p.nq_pad + ql] = li[(t * L + e) * NT + tid];
        }
    }
}

template <int D, int TQ, int TR, bool PACKED>
struct FilterInst {
    static size_t smem_bytes(int list_len) {
        return (size_t)FILTER_STAGES * (D + 1) * TILE_REFS * sizeof(float) + FILTER_STAGES * sizeof(uint64_t) +
               (size_t)TQ * list_len * FILTER_THREADS * (sizeof(float) + sizeof(int));
    }
    static int query(int list_len, int* qpc, int* smem, int* ctas) {
        const size_t sm = smem_bytes(list_len);
        if (sm > 227 * 1024) return -1;
        auto kern = knn_filter_kernel<D, TQ, TR, PACKED>;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess) return -1;
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, FILTER_THREADS, sm) != cudaSuccess) return -1;
        *qpc = FILTER_THREADS * TQ;
        *smem = (int)sm;
        *ctas = nb;
        return nb > 0 ? 0 : -1;
    }
    static int launch(const FilterParams& p, cudaStream_t st) {
        const size_t sm = smem_bytes(p.list_len);
        auto kern = knn_filter_kernel<D, TQ, TR, PACKED>;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess) return -1;
        const unsigned grid = (unsigned)p.nqblocks * (unsigned)p.ref_chunks;
        kern<<<grid, FILTER_THREADS, sm, st>>>(p);
        return cudaGetLastError() == cudaSuccess ? 0 : -1;
    }
};

// op: 0 = query, 1 = launch.  One translation unit per dimension group defines these.
typedef int (*filter_dispatch_fn)(int op, int d, const FilterConfig& cfg, const FilterParams* p, int list_len,
                                  cudaStream_t st, int* qpc, int* smem, int* ctas);
int filter_dispatch_lo(int op, int d, const FilterConfig& cfg, const FilterParams* p, int list_len, cudaStream_t st,
                       int* qpc, int* smem, int* ctas);   // d = 1..16
int filter_dispatch_mid(int op, int d, const FilterConfig& cfg, const FilterParams* p, int list_len, cudaStream_t st,
                        int* qpc, int* smem, int* ctas);  // d = 17..32
int filter_dispatch_hi(int op, int d, const FilterConfig& cfg, const FilterParams* p, int list_len, cudaStream_t st,
                       int* qpc, int* smem, int* ctas);   // d = 33..64

#define MCE_FILTER_CASE(DD, TTQ, TTR, PK)                                                        \
    if (d == DD && cfg.tq == TTQ && cfg.tr == TTR && cfg.packed == PK) {                         \
        if (op == 0) return FilterInst<DD, TTQ, TTR, (PK != 0)>::query(list_len, qpc, smem, ctas); \
        return FilterInst<DD, TTQ, TTR, (PK != 0)>::launch(*p, st);                              \
    }

}  // namespace mce
