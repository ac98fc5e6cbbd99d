// K2: tiled brute-force FP32 filter (the dominant kernel).
//
// Replaces the neighbour search of NearestNeighbors.kneighbors (MCEvidence.py:1104; sklearn
// kd_tree / ArgKmin) by an exhaustive scan whose survivors are re-ranked exactly in FP64 (K3).
//
// Work decomposition: one CTA = one (query block, reference range) unit.  A query block is
// FILTER_THREADS * TQ whitened queries; each thread keeps TQ queries (TQ*D floats) in
// registers for the whole unit.  The first `nfull` blocks scan the whole reference set as one
// unit; the remaining blocks are split into `ref_chunks` ranges to keep the last wave short
// (every split restarts the top-L lists, so it is used sparingly; the split is chosen by the
// cost model in api.cu).  The range is streamed as tiles of tile_refs(D) rows through a 4- or
// 8-slot shared-memory ring filled by TMA bulk copies (cp.async.bulk + mbarrier complete_tx;
// one contiguous (D+1)*tile_refs(D)*4-byte block per tile).  MODE >= 1 (default): slots are
// recycled in groups of half the ring behind one __syncthreads per group; MODE 0: per-slot
// "empty" mbarriers (one arrival per warp) and no CTA-wide barrier.
//
// Ranking score: s(q, r) = q.r - |r|^2/2 = (|q|^2 - |q - r|^2) / 2, larger is closer.  The
// accumulator is initialised with h_r = -|r|^2/2 (row 0 of the tile), so one pair costs exactly
// D FMAs.  In the packed variant two references share one FFMA2 (fma.rn.f32x2): the query
// coordinate is a scalar register broadcast to both halves by the instruction, the reference
// pair comes straight out of a broadcast LDS.128.  MODE >= 2 prefetches the first rows of the
// next step into registers half-way through the current one.
//
// Selection: per step each thread takes the maximum of the TR scores of each of its queries
// (FMNMX3) and compares it with that query's threshold (one FSETP per query, or-reduced to a
// single rarely taken branch).  Passing scores are inserted into a per-(thread, query) list of
// list_len entries in shared memory (replace the minimum, new threshold = new minimum).  A
// thread sees every reference of the range for its own queries, so no cross-thread merge is
// needed; ranges are merged by the re-rank.  In auto mode the self pair always passes and is
// dropped by index inside the slow path, so self-exclusion costs nothing in the fast path.
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
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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

// Ring depth: 8 tiles while they fit 48 KB, else 4.  In SYNC mode the CTA barrier is paid once
// per group of NS/2 tiles (the group just read is refilled while the other group is consumed).
__host__ __device__ constexpr int filter_stages(int d, int tq) {
    return (tq <= 2 && (d + 1) * tile_refs(d) * 4 * 8 <= 48 * 1024) ? 8 : FILTER_STAGES;   // wider register tiles keep occupancy instead
}

// MODE 0: free-running warps (per-slot empty mbarriers); 1: CTA barrier per tile group;
// 2..4: as 1, plus row h and the first MODE-1 coordinate rows of the next step are prefetched
// into registers half-way through the current step (hides the LDS latency at the step start)
template <int D, int TQ, int TR, bool PACKED, int MODE>
__global__ void __launch_bounds__(FILTER_THREADS) knn_filter_kernel(const FilterParams p) {
    constexpr bool SYNC = MODE != 0;
    constexpr bool PF = MODE >= 2 && PACKED;
    constexpr int PFD = PF ? (MODE - 1 < D ? MODE - 1 : D) : 1;   // prefetched coordinate rows
    constexpr int NT = FILTER_THREADS;
    constexpr int TJ = tile_refs(D);
    constexpr int NS = filter_stages(D, TQ);
    constexpr int G = NS / 2;   // tiles per barrier group (SYNC mode)
    constexpr int TILE_FLOATS = (D + 1) * TJ;
    constexpr uint32_t TILE_BYTES = TILE_FLOATS * sizeof(float);
    static_assert(TR == 4 || TR == 8 || TR == 16, "TR must be 4, 8 or 16");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stage = reinterpret_cast<float*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NS * TILE_BYTES);   // full[NS], empty[NS]
    float* ls = reinterpret_cast<float*>(bars + 2 * NS);
    const int L = p.list_len;
    int* li = reinterpret_cast<int*>(ls + TQ * L * NT);

    const int tid = threadIdx.x;
    if (*p.status != 0) return;   // the whitening matrix is garbage: the host reports the error
    // unit -> (query block, reference chunk): the first nfull blocks are whole-set units, the
    // remaining blocks are split so that the last wave of the grid is short
    int qblock, rchunk, t0, t1;
    if ((int)blockIdx.x < p.nfull) {
        qblock = blockIdx.x;
        rchunk = 0;
        t0 = 0;
        t1 = p.ntiles;
    } else {
        const int v = blockIdx.x - p.nfull;
        const int rem = p.nqblocks - p.nfull;
        rchunk = v / rem + p.chunk_base;
        qblock = p.nfull + v % rem;
        chunk_tile_range(p, rchunk, &t0, &t1);
    }
    const long long ql0 = (long long)qblock * (NT * TQ) + tid;  // local id of this thread's first query
    const long long nq = p.q_end - p.q_begin;

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
        if (valid && p.seed_th && rchunk > 0)   // sweep: floor = final threshold of the previous chunk
            th[t] = p.cand_th[cand_th_offset(rchunk - 1, ql, p.nq_pad, p.ntail_pad, p.tail_q0)];
        for (int e = 0; e < L; ++e) {
            ls[(t * L + e) * NT + tid] = -CUDART_INF_F;
            li[(t * L + e) * NT + tid] = -1;
        }
    }
    const int self0 = p.self_exclude ? (int)(p.q_begin + ql0 + p.self_offset) : -0x7fffffff;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            mbar_init(smem_u32(&bars[s]), 1);               // full: one arrive.expect_tx + the TMA bytes
            mbar_init(smem_u32(&bars[NS + s]), NT / 32);    // empty: one arrive per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const float* gt = p.tiles + (long long)t0 * TILE_FLOATS;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < (SYNC ? G : NS - 2); ++s) {
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
        if constexpr (SYNC) {
            if (it % G == 0) {
                // group boundary: every warp is done with the previous group's slots; refill them
                // with the next group while this group (already in flight) is consumed
                if (it > 0) __syncthreads();
                if (tid == 0) {
#pragma unroll
                    for (int i = 0; i < G; ++i) {
                        const int nxt = it + G + i;
                        if (t0 + nxt < t1) {
                            const int ns = nxt % NS;
                            const uint32_t bar = smem_u32(&bars[ns]);
                            mbar_arrive_expect_tx(bar, TILE_BYTES);
                            tma_bulk_g2s(smem_u32(stage + ns * TILE_FLOATS), gt + (long long)nxt * TILE_FLOATS, TILE_BYTES,
                                         bar);
                        }
                    }
                }
            }
        }
        if constexpr (!SYNC) {
            if (tid == 0) {
                // refill two tiles ahead: the slot was read during iteration it-2, so the wait on its
                // "empty" barrier (one arrival per warp) is normally already satisfied and no warp
                // waits for another one at a tile boundary
                const int nxt = it + NS - 2;
                if (t0 + nxt < t1) {
                    const int ns = nxt % NS;
                    if (nxt >= NS) {
                        const uint32_t ebar = smem_u32(&bars[NS + ns]);
                        const uint32_t eparity = ((nxt / NS) - 1) & 1;
                        while (!mbar_try_wait(ebar, eparity)) {
                        }
                    }
                    const uint32_t bar = smem_u32(&bars[ns]);
                    mbar_arrive_expect_tx(bar, TILE_BYTES);
                    tma_bulk_g2s(smem_u32(stage + ns * TILE_FLOATS), gt + (long long)nxt * TILE_FLOATS, TILE_BYTES, bar);
                }
            }
        }
        {
            const uint32_t bar = smem_u32(&bars[slot]);
            while (!mbar_try_wait(bar, parity)) {
            }
        }
        const float* tile = stage + slot * TILE_FLOATS;
        const int jbase = (t0 + it) * TJ;

        float4 pf_r[PFD][TR / 4], pf_h[TR / 4];   // PF: rows h and c < PFD of the step about to start
        if constexpr (PF) {
#pragma unroll
            for (int g = 0; g < TR / 4; ++g) {
#pragma unroll
                for (int cc = 0; cc < PFD; ++cc) pf_r[cc][g] = *reinterpret_cast<const float4*>(tile + (cc + 1) * TJ + 4 * g);
                pf_h[g] = *reinterpret_cast<const float4*>(tile + 4 * g);
            }
        }
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
                        float4 r4;
                        if (PF && c < PFD) r4 = pf_r[c][g];
                        else r4 = *reinterpret_cast<const float4*>(tile + (c + 1) * TJ + j0 + 4 * g);
                        rv[2 * g] = make_float2(r4.x, r4.y);
                        rv[2 * g + 1] = make_float2(r4.z, r4.w);
                    }
                    if (c == 0) {
                        float2 hv[TR / 2];
#pragma unroll
                        for (int g = 0; g < TR / 4; ++g) {
                            float4 h4;
                            if (PF) h4 = pf_h[g];
                            else h4 = *reinterpret_cast<const float4*>(tile + j0 + 4 * g);
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
                    if (PF && c == (D / 2 >= PFD ? D / 2 : D - 1) && (PFD + 1 <= D || j0 + TR < TJ)) {
                        // next step's first rows.  At the last step of a tile the addresses run into the
                        // following rows of the same tile (h is row 0, so row PFD+1 <= D still exists);
                        // those values are never used -- the first step is loaded at the tile start.
#pragma unroll
                        for (int g = 0; g < TR / 4; ++g) {
#pragma unroll
                            for (int cc = 0; cc < PFD; ++cc)
                                pf_r[cc][g] = *reinterpret_cast<const float4*>(tile + (cc + 1) * TJ + j0 + TR + 4 * g);
                            pf_h[g] = *reinterpret_cast<const float4*>(tile + j0 + TR + 4 * g);
                        }
                    }
                }
                float best[TQ];
#pragma unroll
                for (int t = 0; t < TQ; ++t) {
                    best[t] = fmaxf(acc[t][0].x, acc[t][0].y);
#pragma unroll
                    for (int pp = 1; pp < TR / 2; ++pp) best[t] = fmaxf(best[t], fmaxf(acc[t][pp].x, acc[t][pp].y));
                    any |= best[t] > th[t];
                }
                if (any) {
#pragma unroll
                    for (int t = 0; t < TQ; ++t) {
                        if (!(best[t] > th[t])) continue;
#pragma unroll
                        for (int e = 0; e < TR; ++e) {
                            const float s = (e & 1) ? acc[t][e / 2].y : acc[t][e / 2].x;
                            if (s > th[t]) {
                                const int ridx = jbase + j0 + e;
                                if (ridx != self0 + t * NT)
                                    th[t] = fmaxf(th[t], list_insert(ls + t * L * NT + tid, li + t * L * NT + tid, L, s, ridx));
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
                        const float4 r4 = *reinterpret_cast<const float4*>(tile + (c + 1) * TJ + j0 + 4 * g);
                        rv[4 * g] = r4.x;
                        rv[4 * g + 1] = r4.y;
                        rv[4 * g + 2] = r4.z;
                        rv[4 * g + 3] = r4.w;
                    }
                    if (c == 0) {
                        float hv[TR];
#pragma unroll
                        for (int g = 0; g < TR / 4; ++g) {
                            const float4 h4 = *reinterpret_cast<const float4*>(tile + j0 + 4 * g);
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
                float best[TQ];
#pragma unroll
                for (int t = 0; t < TQ; ++t) {
                    best[t] = acc[t][0];
#pragma unroll
                    for (int e = 1; e < TR; ++e) best[t] = fmaxf(best[t], acc[t][e]);
                    any |= best[t] > th[t];
                }
                if (any) {
#pragma unroll
                    for (int t = 0; t < TQ; ++t) {
                        if (!(best[t] > th[t])) continue;
#pragma unroll
                        for (int e = 0; e < TR; ++e) {
                            if (acc[t][e] > th[t]) {
                                const int ridx = jbase + j0 + e;
                                if (ridx != self0 + t * NT)
                                    th[t] = fmaxf(th[t], list_insert(ls + t * L * NT + tid, li + t * L * NT + tid, L, acc[t][e], ridx));
                            }
                        }
                    }
                }
            }
        }
        if constexpr (!SYNC) {
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(smem_u32(&bars[NS + slot]));   // this warp is done with the slot
        }
    }

    // write the lists and final thresholds (coalesced over tid)
#pragma unroll
    for (int t = 0; t < TQ; ++t) {
        const long long ql = ql0 + (long long)t * NT;
        if (ql < nq) {
            p.cand_th[cand_th_offset(rchunk, ql, p.nq_pad, p.ntail_pad, p.tail_q0)] = th[t];
            for (int e = 0; e < L; ++e)
                p.cand_idx[cand_idx_offset(rchunk, e, ql, L, p.nq_pad, p.ntail_pad, p.tail_q0)] = li[(t * L + e) * NT + tid];
        }
    }
}

template <int D, int TQ, int TR, bool PACKED, int MODE>
struct FilterInst {
    static size_t smem_bytes(int list_len) {
        return (size_t)filter_stages(D, TQ) * (D + 1) * tile_refs(D) * sizeof(float) + 2 * filter_stages(D, TQ) * sizeof(uint64_t) +
               (size_t)TQ * list_len * FILTER_THREADS * (sizeof(float) + sizeof(int));
    }
    static int query(int list_len, int* qpc, int* smem, int* ctas) {
        const size_t sm = smem_bytes(list_len);
        if (sm > 227 * 1024) return -1;
        auto kern = knn_filter_kernel<D, TQ, TR, PACKED, MODE>;
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
        auto kern = knn_filter_kernel<D, TQ, TR, PACKED, MODE>;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess) return -1;
        const unsigned grid = (unsigned)p.nfull + (unsigned)(p.nqblocks - p.nfull) * (unsigned)p.ref_chunks;
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

#define MCE_FILTER_CASE1(DD, TTQ, TTR, PK, SY)                                                         \
    if (d == DD && cfg.tq == TTQ && cfg.tr == TTR && cfg.packed == PK && cfg.sync == SY) {             \
        if (op == 0) return FilterInst<DD, TTQ, TTR, (PK != 0), SY>::query(list_len, qpc, smem, ctas); \
        return FilterInst<DD, TTQ, TTR, (PK != 0), SY>::launch(*p, st);                                \
    }

}  // namespace mce
