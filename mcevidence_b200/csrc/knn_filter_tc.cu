// K2-TC: tensor-core engine of the neighbour filter (tcgen05 + TMEM + TMA, sm_100a).  Validated on B200 (bit-identical
// neighbours and logE: tests/test_gpu_tc.py and the whole GPU suite) and chosen automatically for ndim >= 4 from ~3e8 (ndim <= 6: 2e8)
// pairs on ("filter_engine" = 2, the default; 1 forces it, 0 forces the FFMA filter of knn_filter.cuh).
//
// Replaces the same reference step as the FFMA filter, the neighbour search of
// NearestNeighbors.kneighbors (MCEvidence.py:1104), and writes the same candidate planes, so the
// FP64 re-rank / verification (K3) is shared.
//
// One CTA = one (query block of 256, reference range) unit, 20 warps:
//   warps 0-15 epilogue: warp w owns TMEM lane quadrant w%4 of query block (w/4)%2 and column half w/8;
//              thread = one query, 64 columns of every tile
//   warp  16   TMA producer: one bulk copy per 128-reference tile into a shared-memory ring
//   warps 17-18 MMA issuers (even / odd tiles): per tile 2 query blocks x K/8 tcgen05.mma.kind::tf32 (M=128 queries x
//              N=128 references) into one of two TMEM accumulator buffers, one tcgen05.commit per tile
//   warp  19   selection: pops the hits from a shared-memory queue, evaluates them exactly, owns the top-L lists
// A row = [q_0..q_{d-1}, 1, 1, 0..], B row = [r_0..r_{d-1}, h_hi, h_lo, 0..] with h = -|r|^2/2, so
// the accumulator holds the ranking score s = q.r - |r|^2/2 of the FFMA filter, evaluated with
// TF32-truncated operands ("coarse").  Both operand tiles keep the full FP32 words; the tensor
// core ignores the low 13 mantissa bits.
//
// Selection.  An epilogue thread reads its 64 accumulator columns with one tcgen05.ld, takes the maximum of each
// 32-column group (FMNMX3 tree, about half an ALU instruction per pair) and compares it with th - e, where th is
// the query's current threshold (minimum of its top-L list) and e bounds |coarse - exact| for
// every reference (operand truncation 2^-10 each, h split into a TF32-exact head and a remainder,
// generous allowance for the accumulation inside the tensor core).  Only columns that pass are
// evaluated exactly: the FP32 FMA chain of the FFMA filter, read from the very same operand tiles,
// and inserted if they beat th.  Hence no reference whose FP32 score exceeds the final threshold
// is ever dropped, which is the invariant K3's verification relies on.
//
// Hits.  A query meets ~L (1 + ln T) references that beat its running threshold over T tiles, most of them early, and
// what a hit costs is the latency of its path in the owning warp (the sixteen epilogue warps run in lockstep).  Hence
//   (1) long units first look at their first min(128, nt / 12) tiles for the threshold alone (group maxima through an
//       insertion network, nothing evaluated exactly), then scan all tiles from a warm threshold;
//   (2) during the first 300 L tiles a hit lane finds its passing columns itself and pushes (query, reference) items
//       into a shared-memory queue; afterwards it only spills the 32 coarse scores of the group into its slab (stores
//       and one reduction) and the selection warp finds the columns.
// DESIGN.md (K2-TC) has the measurements behind both, and those of the variants that lost.
#include "common.cuh"
#include "knn_filter.cuh"
#include "tcgen05.cuh"
#include <math_constants.h>
#include <type_traits>

// Development switches: variants are built with -D (tools/build_variant.sh) and compared on one GPU box (tools/tc_ab.py).
#ifndef TC_SLAB
#define TC_SLAB 1           // 1: after the start-up tiles a hit lane only spills its 32 coarse scores (fire and forget)
#endif
#ifndef TC_SLAB_T0
#define TC_SLAB_T0 300      // x list length = tiles of a unit that keep the column-precise hit path: about 350 x L / t hits per
#endif                      // tile reach the selection warp, which handles ~2 slab events per tile time (profiles/r02_tc_ab_*.jsonl)
#ifndef TC_BOOT_TILES
#define TC_BOOT_TILES 128   // tiles of a (long) unit that are first looked at for the threshold only (even; 0 = first tile only)
#endif
#ifndef TC_BOOT_DIV
#define TC_BOOT_DIV 12      // ... over nt / 12 tiles at most (nt / 6 for units shorter than 128 tiles)
#endif
#ifndef TC_SEL_MIN
#define TC_SEL_MIN 32       // the selection warp evaluates a batch when at least this many hits are waiting
#endif
#ifndef TC_PROBE
#define TC_PROBE 0          // 1: count hits and the clocks spent in the hit path / in the selection warp (mce_tc_probe)
#endif

namespace mce {

namespace {

constexpr int TC_EPI_WARPS = 16;             // warp w: TMEM lane quadrant w%4, query block (w/4)%2, column half w/8
constexpr int TC_PRODUCER_WARP = 16;
constexpr int TC_MMA_WARP = 17;              // issues the MMAs of the even tiles (of every tile with one issuing warp)
constexpr int TC_MMA_WARP2 = 18;             // issues the MMAs of the odd tiles when tp.mma_warps == 2
constexpr int TC_SELECT_WARP = 19;           // exact evaluation of the hits, owner of the top-L lists
constexpr int TC_HALF_COLS = TC_TILE_REFS / 2;   // columns of a tile looked at by one epilogue thread
constexpr int TC_THREADS = (TC_EPI_WARPS + 4) * 32;
constexpr int TC_QBLK = 128;                 // UMMA M
constexpr uint32_t TC_TMEM_COLS = 512;       // 2 accumulator buffers x 2 query blocks x 128 columns
constexpr uint32_t TC_CHUNK_BYTES = (TC_TILE_REFS / 8) * 128;   // one 16-byte K chunk of all rows of a tile: 2048
constexpr uint32_t TC_ROWGROUP_BYTES = 128;                     // one 8-row core matrix
constexpr int TC_CHUNK_FLOATS = TC_CHUNK_BYTES / 4;             // 512
constexpr int TC_SMEM_BUDGET = 200 * 1024;
// instruction descriptor (kind::tf32): D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), both K-major
// (bits 15, 16 = 0), N >> 3 at bits 17-22, M >> 4 at bits 24-28
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_TILE_REFS >> 3) << 17) |
                              ((uint32_t)(TC_QBLK >> 4) << 24);

// maximum of 32 scores in three levels of FMNMX(3), 17 instructions
__device__ __forceinline__ float max32(const uint32_t* v) {
    float a[11];
#pragma unroll
    for (int j = 0; j < 10; ++j)
        a[j] = max3(__uint_as_float(v[3 * j]), __uint_as_float(v[3 * j + 1]), __uint_as_float(v[3 * j + 2]));
    a[10] = fmaxf(__uint_as_float(v[30]), __uint_as_float(v[31]));
    const float b0 = max3(a[0], a[1], a[2]), b1 = max3(a[3], a[4], a[5]), b2 = max3(a[6], a[7], a[8]);
    const float b3 = fmaxf(a[9], a[10]);
    return fmaxf(max3(b0, b1, b2), b3);
}

// Shared memory is addressed explicitly (32-bit shared-window addresses).
__device__ __forceinline__ float lds_f32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float lds_f32_volatile(uint32_t a) {
    float v;
    asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds_u32_volatile(uint32_t a) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long lds_u64_volatile(uint32_t a) {
    unsigned long long v;
    asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_s32(uint32_t a, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u32_volatile(uint32_t a, uint32_t v) {
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_u64_volatile(uint32_t a, unsigned long long v) {
    asm volatile("st.volatile.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t atoms_add_u32(uint32_t a, uint32_t v) {
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void sts_b32x4(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void reds_or_u32(uint32_t a, uint32_t v) { asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void reds_and_u32(uint32_t a, uint32_t v) { asm volatile("red.shared.and.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ float4 lds_f32x4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}

constexpr uint32_t TC_LIST_STRIDE = TC_QUERIES * 4;   // bytes between consecutive entries of one query's list
constexpr int TC_LMAX = MCE_TC_LIST_MAX;              // longest list: KCAP neighbours + margin
constexpr int TC_BOOT = 9;                            // bootstrap network: the 9 largest coarse scores of a unit's first tile
#ifndef TC_BOOT_DEEP_N
#define TC_BOOT_DEEP_N 13
#endif
constexpr int TC_BOOT_DEEP = TC_BOOT_DEEP_N;                      // multi-tile bootstrap of long lists (kmax 8 .. 10): 13 largest group maxima
constexpr uint32_t TC_QCAP = 512;                     // hit-queue slots (power of two)
constexpr unsigned long long TC_QEMPTY = 0xFFFFFFFFFFFFFFFFull;
// TC_PROBE builds: [0] hit-path entries (lanes), [1] clocks inside, [2] entries that found their slab taken, [3] clocks of
// the column-precise path, [4] its entries, [5] batches evaluated, [6] items in them, [7] clocks of the evaluations,
// [8] slab events, [9] clocks of the slab gathering
static __device__ unsigned long long g_tc_probe[64];   // [16..39] / [40..63]: log2 histograms of the slab / scan path clocks
#if TC_PROBE
#define TC_PROBE_ADD(i, v) atomicAdd(&g_tc_probe[i], (unsigned long long)(v))
#else
#define TC_PROBE_ADD(i, v) do { } while (0)
#endif
// Slab path (TC_SLAB): one 128-byte slab per epilogue thread for the 32 coarse scores of a passing column group, a header
// word per thread (0 = slab free, else 0x80000000 | index of the group's first reference) and one "pending lanes" word
// per epilogue warp.  Only for K <= 40 (ndim <= 38): beyond that the operand tiles need the shared memory.
constexpr int TC_SLAB_MAX_KSTEPS = 5;
constexpr size_t TC_SLAB_BYTES = (size_t)TC_EPI_WARPS * 32 * (128 + 4) + TC_EPI_WARPS * 4 + TC_QUERIES * 4 + 96 * 8;
__host__ __device__ constexpr bool tc_slab_on(int ksteps) { return TC_SLAB != 0 && ksteps <= TC_SLAB_MAX_KSTEPS; }

// Replace the minimum of the list of one query; the caller guarantees s > current minimum.  Returns the new
// minimum = new threshold.  All entries are loaded before the first comparison (one shared-memory latency).
template <int LM>
__device__ __forceinline__ float tc_list_insert(uint32_t ls, uint32_t li, int len, float s, int idx) {
    float v[LM];
#pragma unroll
    for (int e = 0; e < LM; ++e) v[e] = e < len ? lds_f32(ls + e * TC_LIST_STRIDE) : CUDART_INF_F;
    float m = v[0];
#pragma unroll
    for (int e = 1; e < LM; ++e) m = fminf(m, v[e]);
    int am = 0;
#pragma unroll
    for (int e = LM - 1; e >= 0; --e) am = (v[e] == m) ? e : am;   // first entry holding the minimum
    float second = CUDART_INF_F;
#pragma unroll
    for (int e = 0; e < LM; ++e) second = fminf(second, e == am ? CUDART_INF_F : v[e]);
    sts_f32(ls + am * TC_LIST_STRIDE, s);
    sts_s32(li + am * TC_LIST_STRIDE, idx);
    return fminf(s, second);
}

// Exact FP32 score: the FMA chain of the FFMA filter (accumulator initialised with h = h_hi + h_lo, coordinates in
// order) on the FP32 words of the two operand rows.  Element k of a row sits (k >> 2) chunks after the row's first
// 16-byte chunk.  Both rows are loaded completely before the chain starts.
template <int KSTEPS>
__device__ __forceinline__ float tc_score_chain(const float (&a)[8 * KSTEPS], const float (&b)[8 * KSTEPS], int d) {
    float h_hi = 0.f, h_lo = 0.f;
#pragma unroll
    for (int k = 1; k < 8 * KSTEPS; ++k) {
        h_hi = (k == d) ? b[k] : h_hi;
        h_lo = (k == d + 1) ? b[k] : h_lo;
    }
    float acc = h_hi + h_lo;          // == h exactly (h_lo = h - h_hi is exact)
#pragma unroll
    for (int k = 0; k < 8 * KSTEPS; ++k)
        if (k < d) acc = fmaf(a[k], b[k], acc);
    return acc;
}
// both rows in shared memory (validation builds: the operand tile is still in its ring slot)
template <int KSTEPS>
__device__ __forceinline__ float tc_exact_score(uint32_t arow, uint32_t bcol, int d) {
    float a[8 * KSTEPS], b[8 * KSTEPS];
#pragma unroll
    for (int c = 0; c < 2 * KSTEPS; ++c) {
        const float4 av = lds_f32x4(arow + c * TC_CHUNK_BYTES);
        const float4 bv = lds_f32x4(bcol + c * TC_CHUNK_BYTES);
        a[4 * c] = av.x; a[4 * c + 1] = av.y; a[4 * c + 2] = av.z; a[4 * c + 3] = av.w;
        b[4 * c] = bv.x; b[4 * c + 1] = bv.y; b[4 * c + 2] = bv.z; b[4 * c + 3] = bv.w;
    }
    return tc_score_chain<KSTEPS>(a, b, d);
}
// query row in shared memory, reference row from the image in global memory (L2): the selection warp works on
// references whose tile has long left the ring.  Four 16-byte chunks of both rows at a time.
template <int KSTEPS>
__device__ __forceinline__ float tc_exact_score_g(uint32_t arow, const float* __restrict__ brow, int d) {
    constexpr int NC = 2 * KSTEPS;
    float acc = __ldg(brow + (d >> 2) * TC_CHUNK_FLOATS + (d & 3)) +
                __ldg(brow + ((d + 1) >> 2) * TC_CHUNK_FLOATS + ((d + 1) & 3));   // h_hi + h_lo == h exactly
#pragma unroll
    for (int c0 = 0; c0 < NC; c0 += 4) {
        float4 bv[4], av[4];
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (c0 + c < NC) bv[c] = __ldg(reinterpret_cast<const float4*>(brow + (c0 + c) * TC_CHUNK_FLOATS));
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (c0 + c < NC) av[c] = lds_f32x4(arow + (c0 + c) * TC_CHUNK_BYTES);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c0 + c < NC) {
                const int k = 4 * (c0 + c);
                if (k < d) acc = fmaf(av[c].x, bv[c].x, acc);
                if (k + 1 < d) acc = fmaf(av[c].y, bv[c].y, acc);
                if (k + 2 < d) acc = fmaf(av[c].z, bv[c].z, acc);
                if (k + 3 < d) acc = fmaf(av[c].w, bv[c].w, acc);
            }
        }
    }
    return acc;
}

// Two (query, reference) pairs per lane: the loads of both reference rows are issued before the first FMA, so the pair
// shares one L2 round trip (selection warp, batches of 64 hits).
template <int KSTEPS>
__device__ __forceinline__ void tc_exact_score_g2(const uint32_t (&arow)[2], const float* const (&brow)[2], int d, float (&acc)[2]) {
    constexpr int NC = 2 * KSTEPS;
    constexpr int CH = NC < 4 ? NC : 4;
#pragma unroll
    for (int j = 0; j < 2; ++j)
        acc[j] = __ldg(brow[j] + (d >> 2) * TC_CHUNK_FLOATS + (d & 3)) +
                 __ldg(brow[j] + ((d + 1) >> 2) * TC_CHUNK_FLOATS + ((d + 1) & 3));   // h_hi + h_lo == h exactly
#pragma unroll
    for (int c0 = 0; c0 < NC; c0 += CH) {
        float4 bv[2][CH];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int c = 0; c < CH; ++c)
                if (c0 + c < NC) bv[j][c] = __ldg(reinterpret_cast<const float4*>(brow[j] + (c0 + c) * TC_CHUNK_FLOATS));
#pragma unroll
        for (int j = 0; j < 2; ++j) {
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                if (c0 + c < NC) {
                    const float4 av = lds_f32x4(arow[j] + (c0 + c) * TC_CHUNK_BYTES);
                    const int k = 4 * (c0 + c);
                    if (k < d) acc[j] = fmaf(av.x, bv[j][c].x, acc[j]);
                    if (k + 1 < d) acc[j] = fmaf(av.y, bv[j][c].y, acc[j]);
                    if (k + 2 < d) acc[j] = fmaf(av.z, bv[j][c].z, acc[j]);
                    if (k + 3 < d) acc[j] = fmaf(av.w, bv[j][c].w, acc[j]);
                }
            }
        }
    }
}

// hit queue: epilogue lanes push (query, reference), the selection warp pops.  A slot holds TC_QEMPTY or an item, so
// the slot itself says when its content is valid; `head` is published by the consumer after it has emptied slots.
struct TcQueue {
    uint32_t ctrl;    // shared address of {tail, head, done, pad}
    uint32_t items;   // shared address of TC_QCAP 64-bit slots
};
// The store sits INSIDE the wait loop: a lane whose position is within capacity must write its item before it can be
// held at a reconvergence point by warp-mates that are still waiting for room (the consumer cannot pass an unwritten
// slot, and those mates wait for the consumer).
__device__ __forceinline__ void tc_push(const TcQueue& q, uint32_t query, uint32_t ridx) {
    const uint32_t pos = atoms_add_u32(q.ctrl, 1u);
    const unsigned long long item = ((unsigned long long)query << 32) | (unsigned long long)ridx;
    const uint32_t addr = q.items + (pos & (TC_QCAP - 1)) * 8u;
    long long t0 = 0;
    for (unsigned int it = 1u;; ++it) {
        if (pos - lds_u32_volatile(q.ctrl + 4) < TC_QCAP) {
            sts_u64_volatile(addr, item);
            break;
        }
        __nanosleep(64);
        if ((it & 255u) == 0u) {   // the give-up bookkeeping costs an L2 round trip: not on every poll
            if (*reinterpret_cast<volatile unsigned int*>(&g_tc_stuck[0]) != 0u) break;
            const long long now = clock64();
            if (t0 == 0) {
                t0 = now;
            } else if (now - t0 > 1000000000LL) {
                if (atomicCAS(&g_tc_stuck[0], 0u, 2u) == 0u) {
                    g_tc_stuck[1] = pos;
                    g_tc_stuck[2] = lds_u32_volatile(q.ctrl + 4);
                    g_tc_stuck[3] = blockIdx.x;
                }
                break;
            }
        }
    }
}

struct TcEpiState {
    uint32_t th_addr;    // shared address of this query's current threshold (written by the selection warp)
    float e_q;           // bound on |coarse - exact| for this query
    bool valid;
    // validation builds only
    uint32_t arow;       // shared address of this query's row in the A tile
    float worst;         // max |coarse - exact| / e over the pairs of this thread
    unsigned int checked;
    uint32_t full_bar, full_parity;   // "TMA bytes landed" barrier of the tile being processed
    int d;
    float* dump_row;     // where this query's coarse scores of the unit's first tile go (or null)
    uint32_t slab0;      // slab path: shared address of the slab area (slabs, header words, pending-lanes words)
};

// One accumulator half-tile of a query: 64 coarse scores in registers (the TMEM buffer has already been released).
// The maxima of the two 32-column groups (2 x 17 FMNMX) are compared with the query's cut; only then (rare) the
// passing columns are collected in bit masks and pushed to the hit queue.
template <int KSTEPS, bool CHECK, bool USE_SLAB>
__device__ __forceinline__ void tc_process_tile(const uint32_t (&v)[TC_HALF_COLS], uint32_t tile, int col_base, int ridx0, float tcut,
                                                uint32_t query, const TcQueue& q, TcEpiState& st) {
    if constexpr (CHECK) {
        if (st.dump_row) {
#pragma unroll
            for (int i = 0; i < TC_HALF_COLS; ++i) st.dump_row[col_base + i] = __uint_as_float(v[i]);
        }
        if (st.valid) {
            tc_wait(st.full_bar, st.full_parity);   // observe the tile's own TMA barrier before reading the tile
#pragma unroll 1
            for (int i = 0; i < TC_HALF_COLS; ++i) {
                float coarse = 0.f;
#pragma unroll
                for (int j = 0; j < TC_HALF_COLS; ++j)
                    if (j == i) coarse = __uint_as_float(v[j]);
                const int col = col_base + i;
                const float ex = tc_exact_score<KSTEPS>(st.arow, tile + (col >> 3) * TC_ROWGROUP_BYTES + (col & 7) * 16, st.d);
                if (ex > -CUDART_INF_F && st.e_q > 0.f) {
                    st.worst = fmaxf(st.worst, fabsf(coarse - ex) / st.e_q);
                    st.checked++;
                }
            }
        }
    }
    float m[2];
#if defined(TC_HALF_TREE)   // timing experiment: half of the max tree (results are garbage)
    m[0] = max32(&v[0]);
    m[1] = fmaxf(__uint_as_float(v[32]), __uint_as_float(v[63]));
#else
#pragma unroll
    for (int g = 0; g < 2; ++g) m[g] = max32(&v[32 * g]);
#endif
    if (fmaxf(m[0], m[1]) > tcut) {
        // The pipeline has no slack: whatever this warp spends here, all twenty wait for two tiles later (measured: the
        // cost of a hit is the LATENCY of this path, not its instruction count and not the selection warp's time).
#if TC_PROBE
        const long long pc0 = clock64();
#endif
        if constexpr (USE_SLAB) {
            // Stores and a reduction only -- nothing this lane waits for, unless its slab is still taken (rare).  The
            // selection warp finds the passing columns in the slab.
            const uint32_t slab = st.slab0 + threadIdx.x * 128u, hdr = st.slab0 + TC_EPI_WARPS * 32 * 128 + threadIdx.x * 4u;
            const uint32_t pend = st.slab0 + TC_EPI_WARPS * 32 * 132 + (threadIdx.x >> 5) * 4u;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                if (m[g] > tcut) {
                    long long t0 = 0;
#if TC_PROBE
                    if (lds_u32_volatile(hdr) != 0u) TC_PROBE_ADD(2, 1);
#endif
                    for (unsigned int it = 1u; lds_u32_volatile(hdr) != 0u; ++it) {
                        __nanosleep(32);
                        if ((it & 255u) == 0u) {
                            if (*reinterpret_cast<volatile unsigned int*>(&g_tc_stuck[0]) != 0u) break;
                            const long long now = clock64();
                            if (t0 == 0) {
                                t0 = now;
                            } else if (now - t0 > 1000000000LL) {
                                if (atomicCAS(&g_tc_stuck[0], 0u, 3u) == 0u) {
                                    g_tc_stuck[1] = hdr;
                                    g_tc_stuck[2] = threadIdx.x;
                                    g_tc_stuck[3] = blockIdx.x;
                                }
                                break;
                            }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        sts_b32x4(slab + 16u * j, v[32 * g + 4 * j], v[32 * g + 4 * j + 1], v[32 * g + 4 * j + 2], v[32 * g + 4 * j + 3]);
                    sts_u32_volatile(hdr, 0x80000000u | (uint32_t)(ridx0 + col_base + 32 * g));
                    __threadfence_block();
                    reds_or_u32(pend, 1u << (threadIdx.x & 31));
                }
            }
#if TC_PROBE
            TC_PROBE_ADD(0, 1);
            {
                const long long dt = clock64() - pc0;
                TC_PROBE_ADD(1, dt);
                
            }
#endif
        } else {
            uint32_t mask[2];
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                mask[g] = 0u;
                if (m[g] > tcut) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) mask[g] |= (__uint_as_float(v[32 * g + i]) > tcut) ? (1u << i) : 0u;
                }
            }
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                uint32_t mk = mask[g];
                while (mk) {
                    const int col = col_base + 32 * g + (__ffs((int)mk) - 1);
                    mk &= mk - 1u;
                    tc_push(q, query, (uint32_t)(ridx0 + col));
                }
            }
#if TC_PROBE
            TC_PROBE_ADD(4, 1);
            {
                const long long dt = clock64() - pc0;
                TC_PROBE_ADD(3, dt);
                
            }
#endif
        }
    }
}

// Descending insertion network: keeps the NB largest values seen.
template <int NB>
__device__ __forceinline__ void tc_boot_insert(float (&t)[NB], float x) {
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        const float hi = fmaxf(t[i], x);
        x = fminf(t[i], x);
        t[i] = hi;
    }
}

// TL builds record clock64() of the hand-over points of the first 1024 tiles of CTA 0 into tp.timeline ([tile][16]):
// MMA warps 0 tile landed, 1 accumulator free, 2 MMAs + commit issued; epilogue warp 0 (query block 0, columns 0-63)
// 5 accumulator full, 6 accumulator released, 7 tile done; epilogue warp 4 (block 1) 8, 9, 10; producer 11 slot free
// MODE: 0 production, 1 validation (every pair also evaluated exactly, first tile dumped), 2 timeline, 3 / 4 timing
// experiments whose results are garbage (3: the epilogue skips the TMEM loads, 4: loads without the threshold test,
// 5: loads and threshold test but no column ever passes -- the cost of the scan without any hit)
#define TC_MARK(IT, SLOT)                                                                            \
    do {                                                                                             \
        if constexpr (MODE == 2) {                                                                          \
            if (blockIdx.x == 0 && (IT) < 1024 && lane == 0) tp.timeline[(IT) * 16 + (SLOT)] = clock64(); \
        }                                                                                            \
    } while (0)

template <int KSTEPS, int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1) knn_filter_tc_kernel(const TcFilterParams tp) {
    constexpr bool CHECK = MODE == 1;
    constexpr int K = KSTEPS * 8;
    constexpr uint32_t TILE_BYTES = (uint32_t)K * TC_TILE_REFS * 4;   // K/4 chunks of 2 KB
    constexpr uint32_t A_MB_BYTES = (uint32_t)K * TC_QBLK * 4;
    constexpr uint32_t KSTEP_BYTES = 2 * TC_CHUNK_BYTES;              // one MMA consumes K = 8 TF32 = two chunks
    const FilterParams& p = tp.f;
    const int NS = tp.stages;
    const int L = p.list_len;

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float* a_sm = reinterpret_cast<float*>(smem_raw);
    float* ring = reinterpret_cast<float*>(smem_raw + 2 * A_MB_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + 2 * A_MB_BYTES + (size_t)NS * TILE_BYTES);
    uint64_t* bar_full = bars;              // [NS]  TMA bytes of a tile have landed
    uint64_t* bar_free = bars + NS;         // [NS]  all epilogue warps are done with the tile in the slot
    uint64_t* bar_acc_full = bars + 2 * NS; // [2 buffers]  the MMAs of a tile (both query blocks) have completed
    uint64_t* bar_acc_free = bars + 2 * NS + 4;   // [2 buffers]  the eight epilogue warps have read the accumulator
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS + 8);
    uint32_t* qctrl = tmem_slot + 4;        // hit queue: tail, head, done, pad
    unsigned long long* qitems = reinterpret_cast<unsigned long long*>(qctrl + 4);
    float* th_s = reinterpret_cast<float*>(qitems + TC_QCAP);   // [256] current threshold of every query
    float* ls = th_s + TC_QUERIES;                              // [L][256] list scores
    int* li = reinterpret_cast<int*>(ls + (size_t)L * TC_QUERIES);   // [L][256] list indices
    constexpr bool SLAB = tc_slab_on(KSTEPS);
    // slab path: [512] slabs of 32 floats, [512] header words, [16] pending-lanes words, [256] e_q, [96] staged items
    float* slabs = reinterpret_cast<float*>(li + (size_t)L * TC_QUERIES);
    uint32_t* slab_hdr = reinterpret_cast<uint32_t*>(slabs + TC_EPI_WARPS * 32 * 32);
    uint32_t* slab_pend = slab_hdr + TC_EPI_WARPS * 32;
    float* eq_s = reinterpret_cast<float*>(slab_pend + TC_EPI_WARPS);
    unsigned long long* stage = reinterpret_cast<unsigned long long*>(eq_s + TC_QUERIES);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    if (*p.status != 0) return;   // the whitening matrix is garbage: the host reports the error

    // unit -> (query block, reference chunk), as in the FFMA filter
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
    const int nt = t1 - t0;
    const long long nq = p.q_end - p.q_begin;
    // Long units look at their first `boot` tiles twice: once for the threshold alone (the epilogue keeps the largest
    // group maxima, nothing is evaluated exactly), then again as part of the normal pass, which starts from a
    // threshold that has seen boot x 128 references.  Two thirds of the hits of a unit used to fall into these tiles,
    // where the selection warp -- not the tensor core -- set the pace.
    int boot = 0;   // even; a twelfth of the unit (a sixth below 128 tiles), under 4 tiles the first-tile bootstrap of the epilogue
    if (MODE == 0 && TC_BOOT_TILES > 0 && L + 1 <= TC_BOOT_DEEP && !(p.seed_th && rchunk > 0)) {
        boot = (nt / (nt < 128 ? TC_BOOT_DIV / 2 : TC_BOOT_DIV)) & ~1;   // short units (small chains, split references): a sixth
        boot = boot > TC_BOOT_TILES ? TC_BOOT_TILES : (boot < 4 ? 0 : boot);
    }
    const int ntv = nt + boot;   // tiles in the order they are loaded and multiplied: 0 .. boot-1, then 0 .. nt-1

    // ---- set-up ----------------------------------------------------------------------------------
    if (warp == TC_MMA_WARP) tmem_alloc(smem_u32(tmem_slot), TC_TMEM_COLS);
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(smem_u32(&bar_full[s]), 1);
            mbar_init(smem_u32(&bar_free[s]), TC_EPI_WARPS);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(smem_u32(&bar_acc_full[b]), 1);
            mbar_init(smem_u32(&bar_acc_free[b]), TC_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        qctrl[0] = 0u; qctrl[1] = 0u; qctrl[2] = 0u; qctrl[3] = 0u;
    }
    for (uint32_t i = tid; i < TC_QCAP; i += TC_THREADS) qitems[i] = TC_QEMPTY;
    if constexpr (SLAB) {
        if (tid < TC_EPI_WARPS * 32) slab_hdr[tid] = 0u;
        if (tid < TC_EPI_WARPS) slab_pend[tid] = 0u;
    }
    TcQueue queue{smem_u32(qctrl), smem_u32(qitems)};
    TcEpiState st;
    st.worst = 0.f;
    st.checked = 0;
    st.dump_row = nullptr;
    const int qslot = tid & (TC_QUERIES - 1), half = (tid >> 8) & 1;   // epilogue threads: query of the block, column half
    const long long ql = (long long)qblock * TC_QUERIES + qslot;         // local query id
    if (warp < TC_EPI_WARPS) {
        const int mb = qslot >> 7, row = qslot & 127;
        const bool valid = ql < nq;
        const double* src = p.yq + (p.q_begin + (valid ? ql : 0)) * p.d;
        float* arow = a_sm + mb * (K * TC_QBLK) + (row >> 3) * 32 + (row & 7) * 4;
        float qn2 = 0.f;
#pragma unroll
        for (int c = 0; c < K / 4; ++c) {
            float w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int k = 4 * c + i;
                float x = 0.f;
                if (valid) {
                    if (k < p.d) {
                        x = (float)src[k];
                        qn2 = fmaf(x, x, qn2);
                    } else if (k < p.d + 2) {
                        x = 1.f;
                    }
                }
                w[i] = x;
            }
            if (half == 0) *reinterpret_cast<float4*>(arow + c * TC_CHUNK_FLOATS) = make_float4(w[0], w[1], w[2], w[3]);
        }
        const float rmax = __uint_as_float(*tp.rmax_bits);
        const float qn = sqrtf(qn2) * 1.000001f;
        st.arow = smem_u32(arow);
        st.th_addr = smem_u32(th_s + qslot);
        st.e_q = tp.slack_qr * qn * rmax + tp.slack_rr * rmax * rmax;
        st.valid = valid;
        st.d = p.d;
        st.slab0 = smem_u32(slabs);
        if (half == 0) {
            if constexpr (SLAB) eq_s[qslot] = st.e_q;
            // sweep: the final threshold of the previous chunk is a floor for this one (see FilterParams::seed_th)
            th_s[qslot] = (valid && p.seed_th && rchunk > 0)
                              ? p.cand_th[cand_th_offset(rchunk - 1, ql, p.nq_pad, p.ntail_pad, p.tail_q0)]
                              : -CUDART_INF_F;
            for (int e = 0; e < L; ++e) {
                ls[e * TC_QUERIES + qslot] = -CUDART_INF_F;
                li[e * TC_QUERIES + qslot] = -1;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // A tile: generic-proxy stores -> tensor-core reads
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // ---- roles ------------------------------------------------------------------------------------
    if (warp == TC_PRODUCER_WARP) {
        if (elect_one()) {
            const unsigned char* gt = reinterpret_cast<const unsigned char*>(p.tiles) + (size_t)t0 * TILE_BYTES;
            int slot = 0, round = 0;
            for (int n = 0; n < ntv; ++n) {
                if (round > 0) {
                    const uint32_t fb = smem_u32(&bar_free[slot]);
                    tc_wait_relaxed(fb, (uint32_t)(round - 1) & 1u);
                }
                if constexpr (MODE == 2) {
                    if (blockIdx.x == 0 && n < 1024) tp.timeline[n * 16 + 11] = clock64();
                }
                const uint32_t bar = smem_u32(&bar_full[slot]);
                mbar_arrive_expect_tx(bar, TILE_BYTES);
                tma_bulk_g2s(smem_u32(ring) + (uint32_t)slot * TILE_BYTES, gt + (size_t)(n < boot ? n : n - boot) * TILE_BYTES, TILE_BYTES, bar);
                if (++slot == NS) {
                    slot = 0;
                    ++round;
                }
            }
        }
    } else if (warp == TC_MMA_WARP || warp == TC_MMA_WARP2) {
        // One hand-over per tile: wait for the tile and for the accumulator buffer, issue the 2 x KSTEPS MMAs of both
        // query blocks, one commit.  The waits and the issue cost ~400 cycles of one thread per tile, more than the
        // MMAs themselves, so with tp.mma_warps == 2 a second warp issues the odd tiles (buffer 1) independently.
        const int nw = tp.mma_warps == 2 ? 2 : 1;
        const int first = warp == TC_MMA_WARP ? 0 : 1;
        if (warp == TC_MMA_WARP || nw == 2) {
            const uint32_t a_base = smem_u32(a_sm), ring_base = smem_u32(ring);
            // debug bits of swap_offsets: 1 = exchange the two byte offsets, 2 = descriptor version field 0 instead of 1
            const uint32_t lbo = (tp.swap_offsets & 1) ? TC_ROWGROUP_BYTES : TC_CHUNK_BYTES;
            const uint32_t sbo = (tp.swap_offsets & 1) ? TC_CHUNK_BYTES : TC_ROWGROUP_BYTES;
            const uint32_t dver = (tp.swap_offsets & 2) ? 0u : 1u;
            const uint64_t desc_hi = tc_smem_desc(0, lbo, sbo, dver);
#if TC_PROBE
            long long probe_mwait = 0;
#endif
            for (int it = first; it < ntv; it += nw) {
                const int buf = it & 1;
                const int slot = it % NS, round = it / NS;
                tc_wait(smem_u32(&bar_full[slot]), (uint32_t)round & 1u);
                TC_MARK(it, 0);
#if TC_PROBE
                const long long pm0 = clock64();
#endif
                if (it >= 2)   // the epilogue warps have drained this buffer (tile it - 2)
                    tc_wait(smem_u32(&bar_acc_free[buf]), (uint32_t)((it >> 1) - 1) & 1u);
#if TC_PROBE
                probe_mwait += clock64() - pm0;
#endif
                TC_MARK(it, 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t b_addr = ring_base + (uint32_t)slot * TILE_BYTES;
#pragma unroll
                    for (int mb = 0; mb < 2; ++mb) {
                        const uint32_t d_tmem = tmem_base + (uint32_t)buf * 256u + (uint32_t)mb * 128u;
#pragma unroll
                        for (int ks = 0; ks < KSTEPS; ++ks) {
                            const uint64_t adesc = desc_hi | (uint64_t)(((a_base + mb * A_MB_BYTES + ks * KSTEP_BYTES) >> 4) & 0x3FFFu);
                            const uint64_t bdesc = desc_hi | (uint64_t)(((b_addr + ks * KSTEP_BYTES) >> 4) & 0x3FFFu);
                            umma_tf32(d_tmem, adesc, bdesc, TC_IDESC, ks > 0 ? 1u : 0u);
                        }
                    }
                    umma_commit(smem_u32(&bar_acc_full[buf]));
                }
                __syncwarp();
                TC_MARK(it, 2);
            }
#if TC_PROBE
            if (lane == 0) {
                TC_PROBE_ADD(32 + first, probe_mwait);   // [32..33] clocks an MMA warp waited for "accumulator free"
                TC_PROBE_ADD(34, nt);                     // [34] tiles (counted by both MMA warps)
            }
#endif
        }
    } else if (warp == TC_SELECT_WARP) {
        // Selection warp: collects hits -- (query, reference) items pushed by the epilogue lanes during the start-up
        // tiles, column groups spilled to slabs afterwards -- evaluates them exactly in batches of up to 64 (two per
        // lane: query row from the A tile, reference row from the image in L2) and maintains the lists and thresholds
        // of all 256 queries.
        const uint32_t ctrl = queue.ctrl;
        const float* img = p.tiles;
        const int d = p.d;
        // exact evaluation and insertion of up to 64 items, lane l holding items l and 32 + l of the batch
        auto eval_insert = [&](const unsigned long long (&item)[2], const bool (&has)[2], const uint32_t n) {
#if TC_PROBE
            const long long pe0 = clock64();
#endif
            uint32_t q[2];
            int ridx[2];
            float s[2];
            uint32_t arow[2];
            const float* brow[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                q[j] = has[j] ? (uint32_t)(item[j] >> 32) : 0u;
                ridx[j] = has[j] ? (int)(uint32_t)item[j] : 0;
                arow[j] = smem_u32(a_sm) + (q[j] >> 7) * A_MB_BYTES + ((q[j] & 127u) >> 3) * TC_ROWGROUP_BYTES + (q[j] & 7u) * 16u;
                brow[j] = img + (size_t)(ridx[j] >> 7) * (K * TC_TILE_REFS) + ((ridx[j] & 127) >> 3) * 32 + (ridx[j] & 7) * 4;
            }
            tc_exact_score_g2<KSTEPS>(arow, brow, d, s);   // both rows' loads in flight together (idle lanes read row 0)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int self = p.self_exclude ? (int)(p.q_begin + (long long)qblock * TC_QUERIES + q[j] + p.self_offset) : -0x7fffffff;
                if (!has[j] || ridx[j] == self) s[j] = -CUDART_INF_F;
            }
            // hits of the same query are inserted one after the other (rank within the half-batch, first half first)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (j == 1 && n <= 32u) continue;
                const bool hj = has[j];
                const uint32_t qj = q[j];
                const float sj = s[j];
                const int rj = ridx[j];
                const uint32_t peers = __match_any_sync(0xffffffffu, hj ? qj : (0x80000000u | (uint32_t)lane));
                const int rank = __popc(peers & ((1u << lane) - 1u));
                const int maxrank = __reduce_max_sync(0xffffffffu, rank);
                for (int r = 0; r <= maxrank; ++r) {
                    if (hj && rank == r) {
                        const uint32_t tha = smem_u32(th_s) + qj * 4u;
                        const float th = lds_f32(tha);
                        if (sj > th) {
                            const uint32_t lsa = smem_u32(ls) + qj * 4u, lia = smem_u32(li) + qj * 4u;
                            const float nmin = L <= 8 ? tc_list_insert<8>(lsa, lia, L, sj, rj) : tc_list_insert<TC_LMAX>(lsa, lia, L, sj, rj);
                            sts_f32(tha, fmaxf(th, nmin));   // the bootstrap floor stays in force while the list fills
                        }
                    }
                    __syncwarp();
                }
            }
#if TC_PROBE
            if (lane == 0) {
                TC_PROBE_ADD(5, 1);
                TC_PROBE_ADD(6, n);
                TC_PROBE_ADD(7, clock64() - pe0);
            }
#endif
        };
        const uint32_t stage_a = smem_u32(stage), hdr_a = smem_u32(slab_hdr), pend_a = smem_u32(slab_pend), slab_a = smem_u32(slabs);
        const uint32_t th_a = smem_u32(th_s), eq_a = smem_u32(eq_s);
        uint32_t head = 0, idle = 0, staged = 0;   // staged: items found in slabs, waiting in `stage` (this warp only)
        auto flush_stage = [&]() {
            unsigned long long item[2];
            bool has[2];
            __syncwarp();   // the items were staged by other lanes
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                has[j] = 32u * j + (uint32_t)lane < staged;
                item[j] = has[j] ? lds_u64_volatile(stage_a + (32u * j + lane) * 8u) : TC_QEMPTY;
            }
            __syncwarp();
            eval_insert(item, has, staged);
            staged = 0;
        };
        for (;;) {
            const uint32_t done = lds_u32_volatile(ctrl + 8);
            const uint32_t tail = lds_u32_volatile(ctrl);
            const uint32_t avail = tail - head;
            bool progress = false;
            if constexpr (SLAB) {
                // column groups spilled by the epilogue lanes: find the passing columns (one per lane) against the
                // query's cut of NOW, hand the slab back, stage the items
                // Lane w looks at warp w's pending lanes and takes them off the word at once (no round trip): a lane
                // cannot raise its bit again before its slab has been handed back below.
                const uint32_t pend = lane < TC_EPI_WARPS ? lds_u32_volatile(pend_a + 4u * lane) : 0u;
                if (pend != 0u) {
                    reds_and_u32(pend_a + 4u * lane, ~pend);
                    __threadfence_block();
                }
                uint32_t wmask = __ballot_sync(0xffffffffu, pend != 0u);
                while (wmask) {
                    const int w = __ffs((int)wmask) - 1;
                    wmask &= wmask - 1u;
                    uint32_t lm = __shfl_sync(0xffffffffu, pend, w);
                    while (lm) {
#if TC_PROBE
                        const long long pg0 = clock64();
#endif
                        const int l = __ffs((int)lm) - 1;
                        lm &= lm - 1u;
                        const uint32_t t = (uint32_t)(w * 32 + l), qq = t & (TC_QUERIES - 1);
                        const uint32_t hdr = lds_u32_volatile(hdr_a + 4u * t);
                        const float x = lds_f32_volatile(slab_a + 128u * t + 4u * lane);
                        const float cut = lds_f32_volatile(th_a + 4u * qq) - lds_f32(eq_a + 4u * qq);
                        const uint32_t cm = __ballot_sync(0xffffffffu, x > cut);   // every lane has read its score
                        if (lane == 0) sts_u32_volatile(hdr_a + 4u * t, 0u);      // slab back to its owner
                        if ((cm >> lane) & 1u) {
                            const uint32_t pos = staged + (uint32_t)__popc(cm & ((1u << lane) - 1u));
                            sts_u64_volatile(stage_a + pos * 8u, ((unsigned long long)qq << 32) | (unsigned long long)((hdr & 0x7fffffffu) + (uint32_t)lane));
                        }
                        staged += (uint32_t)__popc(cm);
#if TC_PROBE
                        if (lane == 0) {
                            TC_PROBE_ADD(8, 1);
                            TC_PROBE_ADD(9, clock64() - pg0);
                        }
#endif
                        if (staged >= 32u) flush_stage();
                        progress = true;
                    }
                }
            }
            // A batch costs this warp the same few hundred issue slots whether it holds one hit or 64: wait until it is
            // worth it (thresholds lag by the few tiles that takes, which costs next to nothing).
            if (avail >= (done == (uint32_t)TC_EPI_WARPS ? 1u : (uint32_t)TC_SEL_MIN)) {
                // pop the written prefix of the next 64 positions, two per lane (lane l: positions l and 32 + l) -- the exact
                // evaluation waits for one L2 round trip per batch, so a batch of 64 costs little more than one of 32.  A
                // reserved slot may not have been written yet, and its owner may be held up by warp-mates that wait for
                // `head` to move, so never wait for a slot here.
                const uint32_t nres = avail < 64u ? avail : 64u;
                unsigned long long item[2] = {TC_QEMPTY, TC_QEMPTY};
                uint32_t addr[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    addr[j] = queue.items + ((head + 32u * j + lane) & (TC_QCAP - 1)) * 8u;
                    if (32u * j + (uint32_t)lane < nres) item[j] = lds_u64_volatile(addr[j]);
                }
                const uint32_t unw0 = __ballot_sync(0xffffffffu, (uint32_t)lane < nres && item[0] == TC_QEMPTY);
                const uint32_t unw1 = __ballot_sync(0xffffffffu, 32u + (uint32_t)lane < nres && item[1] == TC_QEMPTY);
                const uint32_t n = unw0 ? (uint32_t)(__ffs((int)unw0) - 1) : unw1 ? 32u + (uint32_t)(__ffs((int)unw1) - 1) : nres;
                if (n != 0u) {
                    bool has[2];
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        has[j] = 32u * j + (uint32_t)lane < n;
                        if (has[j]) sts_u64_volatile(addr[j], TC_QEMPTY);
                    }
                    head += n;
                    __threadfence_block();
                    __syncwarp();
                    if (lane == 0) sts_u32_volatile(ctrl + 4, head);
                    eval_insert(item, has, n);
                    progress = true;
                }
            }
            if (progress) {
                idle = 0;
                continue;
            }
            ++idle;
            if (staged != 0u && (idle >= 8u || done == (uint32_t)TC_EPI_WARPS)) {   // nothing else to do: the leftovers
                flush_stage();
                continue;
            }
            if (done == (uint32_t)TC_EPI_WARPS && avail == 0u && staged == 0u) break;   // `done` was read before the queue and the slabs
            if ((idle & 255u) == 0u && *reinterpret_cast<volatile unsigned int*>(&g_tc_stuck[0]) != 0u) break;
            __nanosleep(avail == 0u ? 64 : 256);
        }
    } else {
        const int mb = (warp >> 2) & 1, quad = warp & 3;
        const int col_base = half * TC_HALF_COLS;
        const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(mb * 128 + col_base);
        const uint32_t acc_full0 = smem_u32(&bar_acc_full[0]), acc_free0 = smem_u32(&bar_acc_free[0]);
        const uint32_t slot_bar0 = smem_u32(&bar_free[0]), slot_bar_end = slot_bar0 + (uint32_t)NS * 8u;
        const uint32_t ring0 = smem_u32(ring);
        // a unit that starts from the thresholds of the chunk before it (sweep) has no start-up phase
        const int slab_t0 = (p.seed_th && rchunk > 0) ? 0 : TC_SLAB_T0 * L;   // even
        // invalid queries never pass: th - e_q is NaN (th starts at -inf) or +inf
        const float e_q = (st.valid && MODE != 5) ? st.e_q : -CUDART_INF_F;   // MODE 5: nothing ever passes

        uint32_t parity = 0u, slot_bar = slot_bar0;
        // The threshold can start from c - e, c the (L+1)-th largest coarse score seen: at least L references other than the
        // query itself have an exact score >= c - e (the larger of the two column halves wins).
        auto raise_threshold = [&](const auto& t) {
            constexpr int NB = (int)(sizeof(t) / sizeof(float));
            float c0 = -CUDART_INF_F;
#pragma unroll
            for (int i = 0; i < NB; ++i) c0 = (i == L) ? t[i] : c0;
            if (st.valid && c0 > -CUDART_INF_F) {
                const float th0 = c0 - st.e_q * 1.001f - 1e-30f;
                // float maximum through compare-and-swap on the bit pattern (twice per query and unit)
                uint32_t old = lds_u32_volatile(st.th_addr);
                while (th0 > __uint_as_float(old)) {
                    uint32_t seen;
                    asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(seen) : "r"(st.th_addr), "r"(old), "r"(__float_as_uint(th0)) : "memory");
                    if (seen == old) break;
                    old = seen;
                }
            }
            __threadfence_block();
        };
        // Bootstrap of a long unit: the maxima of this thread's two 32-column groups of each of the first `boot` tiles
        // (distinct references) go through the insertion network; no hit path, nothing for the selection warp.
        auto run_boot = [&](auto depth_c) {
            constexpr int NB = decltype(depth_c)::value;
            float t[NB];
#pragma unroll
            for (int i = 0; i < NB; ++i) t[i] = -CUDART_INF_F;
            auto boot_tile = [&](const int buf) {
                tc_wait(acc_full0 + 8u * buf, parity);
                __syncwarp();
                tc_fence_after();
                uint32_t v[TC_HALF_COLS];
                tmem_ld64_wait(tlane + 256u * buf, v);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(acc_free0 + 8u * buf);
                    mbar_arrive(slot_bar);
                }
                tc_boot_insert(t, max32(&v[0]));
                tc_boot_insert(t, max32(&v[32]));
                slot_bar += 8u;
                if (slot_bar == slot_bar_end) slot_bar = slot_bar0;
            };
#pragma unroll 1
            for (int j = 0; j < boot; j += 2) {
                boot_tile(0);
                boot_tile(1);
                parity ^= 1u;
            }
            raise_threshold(t);
        };
        if (boot > 0) {
            if (L + 1 <= TC_BOOT) run_boot(std::integral_constant<int, TC_BOOT>{});
            else run_boot(std::integral_constant<int, TC_BOOT_DEEP>{});
        } else if (MODE <= 2 && nt > 0 && L + 1 <= TC_BOOT) {
            // Bootstrap of a short unit: the (L+1)-th largest coarse score among this thread's columns of the first tile,
            // so that the first tile yields a handful of hits per query instead of 128.  The tile is processed again
            // by the loop below.
            tc_wait(acc_full0, 0u);
            __syncwarp();
            tc_fence_after();
            float t[TC_BOOT];
#pragma unroll
            for (int i = 0; i < TC_BOOT; ++i) t[i] = -CUDART_INF_F;
#pragma unroll 1
            for (int c = 0; c < TC_HALF_COLS / 32; ++c) {
                uint32_t w[32];
                tmem_ld32(tlane + 32 * c, w);
                tmem_wait_ld(w);
#pragma unroll
                for (int i = 0; i < 32; ++i) tc_boot_insert(t, __uint_as_float(w[i]));
            }
            raise_threshold(t);
        }

        // Tile loop, unrolled over the two accumulator buffers so that every address is loop-invariant.  Per tile and
        // thread: one barrier wait, one TMEM load of this thread's 64 columns, two arrivals (accumulator buffer back
        // to the MMA issuer, ring slot back to the producer), 35 FMNMX and one compare.
#if TC_PROBE
        long long probe_wait = 0;   // clocks this warp waited for "accumulator full"
#endif
        int ridx0 = t0 * TC_TILE_REFS;
        auto one_tile = [&](const int buf, const int it, auto use_slab_c) {
            constexpr bool US = decltype(use_slab_c)::value;
#if TC_PROBE
            const long long pw0 = clock64();
#endif
            tc_wait(acc_full0 + 8u * buf, parity);
            __syncwarp();   // lanes may leave the wait at different times; the TMEM load is warp-collective
#if TC_PROBE
            probe_wait += clock64() - pw0;
#endif
            tc_fence_after();
            if (quad == 0 && half == 0) TC_MARK(it, 5 + 3 * mb);
            const uint32_t taddr = tlane + 256u * buf;
            if constexpr (MODE <= 2 || MODE == 5) {
                const float tcut = lds_f32_volatile(st.th_addr) - e_q;
                uint32_t v[TC_HALF_COLS];
                tmem_ld64_wait(taddr, v);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(acc_free0 + 8u * buf);
                    if (!CHECK) mbar_arrive(slot_bar);
                }
                if (quad == 0 && half == 0) TC_MARK(it, 6 + 3 * mb);
                if constexpr (CHECK) {
                    const int slot = (int)((slot_bar - slot_bar0) >> 3);
                    st.full_bar = smem_u32(&bar_full[slot]);
                    st.full_parity = (uint32_t)(it / NS) & 1u;
                    st.dump_row = (tp.dump && blockIdx.x == 0 && it == 0) ? tp.dump + (size_t)qslot * TC_TILE_REFS : nullptr;
                    tc_process_tile<KSTEPS, CHECK, US>(v, ring0 + (uint32_t)slot * TILE_BYTES, col_base, ridx0, tcut, (uint32_t)qslot, queue, st);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(slot_bar);
                } else {
                    tc_process_tile<KSTEPS, CHECK, US>(v, 0u, col_base, ridx0, tcut, (uint32_t)qslot, queue, st);
                }
            } else {
                if constexpr (MODE == 4) {   // the loads alone
                    uint32_t v[TC_HALF_COLS];
                    tmem_ld64_wait(taddr, v);
                    st.checked ^= v[0] ^ v[63];
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(acc_free0 + 8u * buf);
                    mbar_arrive(slot_bar);
                }
                if (quad == 0 && half == 0) TC_MARK(it, 6 + 3 * mb);
            }
            if (quad == 0 && half == 0) TC_MARK(it, 7 + 3 * mb);
            slot_bar += 8u;
            if (slot_bar == slot_bar_end) slot_bar = slot_bar0;
            ridx0 += TC_TILE_REFS;
        };
        // two copies of the loop: the start-up tiles with the column-precise hit path, the rest with the slab path
        int it = 0;
        const int nt0 = SLAB ? (nt < slab_t0 ? nt : slab_t0) : nt;   // slab_t0 is even
#pragma unroll 1
        for (; it < nt0; it += 2) {
            one_tile(0, it, std::false_type{});
            if (it + 1 < nt) one_tile(1, it + 1, std::false_type{});
            parity ^= 1u;
        }
        if constexpr (SLAB) {
#pragma unroll 1
            for (; it < nt; it += 2) {
                one_tile(0, it, std::true_type{});
                if (it + 1 < nt) one_tile(1, it + 1, std::true_type{});
                parity ^= 1u;
            }
        }
#if TC_PROBE
        if (lane == 0) TC_PROBE_ADD(16 + warp, probe_wait);   // [16..31]
#endif
        // every hit of this warp is in the queue: tell the selection warp
        __threadfence_block();
        __syncwarp();
        if (lane == 0) atoms_add_u32(queue.ctrl + 8, 1u);
        if constexpr (CHECK) {
            if (tp.debug) {
                atomicMax(&tp.debug[0], __float_as_uint(st.worst));   // non-negative floats order like their bits
                atomicAdd(&tp.debug[1], st.checked);
            }
        } else if (MODE == 4 && st.checked == 0x9e3779b9u && tp.debug) {
            tp.debug[1] = st.checked;   // keeps the loads of the timing experiment alive
        }
    }

    // ---- lists out, tear-down ----------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();   // the selection warp has drained the queue: lists and thresholds are final
    if (warp < TC_EPI_WARPS && half == 0 && ql < nq) {
        p.cand_th[cand_th_offset(rchunk, ql, p.nq_pad, p.ntail_pad, p.tail_q0)] = th_s[qslot];
        for (int e = 0; e < L; ++e)
            p.cand_idx[cand_idx_offset(rchunk, e, ql, L, p.nq_pad, p.ntail_pad, p.tail_q0)] = li[e * TC_QUERIES + qslot];
    }
    if (warp == TC_MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TC_TMEM_COLS);
    }
}

size_t tc_fixed_smem(int K, int list_len) {
    return (size_t)2 * K * TC_QBLK * 4 + (size_t)(2 * 8 + 8) * sizeof(uint64_t) + 16 + 16 + (size_t)TC_QCAP * 8 +
           (size_t)TC_QUERIES * 4 + (size_t)list_len * TC_QUERIES * (sizeof(float) + sizeof(int)) +
           (tc_slab_on(K / 8) ? TC_SLAB_BYTES : 0);
}

template <int KSTEPS, int MODE>
int tc_launch_mode(const TcFilterParams& tp, unsigned grid, size_t smem, cudaStream_t st) {
    auto kern = knn_filter_tc_kernel<KSTEPS, MODE>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    kern<<<grid, TC_THREADS, smem, st>>>(tp);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

template <int KSTEPS>
int tc_launch_k(const TcFilterParams& tp, unsigned grid, size_t smem, cudaStream_t st) {
    if constexpr (KSTEPS == 2) {   // diagnostics exist for ndim 7..14 only
        if (tp.timeline) return tc_launch_mode<KSTEPS, 2>(tp, grid, smem, st);
        if (tp.perf_mode == 1) return tc_launch_mode<KSTEPS, 3>(tp, grid, smem, st);
        if (tp.perf_mode == 2) return tc_launch_mode<KSTEPS, 4>(tp, grid, smem, st);
        if (tp.perf_mode == 3) return tc_launch_mode<KSTEPS, 5>(tp, grid, smem, st);
    }
    if (tp.check) return tc_launch_mode<KSTEPS, 1>(tp, grid, smem, st);
    return tc_launch_mode<KSTEPS, 0>(tp, grid, smem, st);
}

// reference image: one thread per reference row
__global__ void __launch_bounds__(128) tc_image_kernel(const double* __restrict__ y64, const float* __restrict__ ffma_tiles,
                                                        long long n_ref, long long n_pad, int d, int K,
                                                        float* __restrict__ image) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_pad) return;
    const bool valid = j < n_ref;
    const int TJ = tile_refs(d);
    float h_hi = -CUDART_INF_F, h_lo = 0.f;
    if (valid) {
        const float h = ffma_tiles[(j / TJ) * (long long)((d + 1) * TJ) + (j % TJ)];   // row 0 of the FFMA tile
        h_hi = __uint_as_float(__float_as_uint(h) & 0xFFFFE000u);                       // exact in TF32
        h_lo = h - h_hi;                                                                 // exact in FP32
        if (!(fabsf(h) < CUDART_INF_F)) {
            h_hi = h;
            h_lo = 0.f;
        }
    }
    const long long tile = j / TC_TILE_REFS;
    const int col = (int)(j - tile * TC_TILE_REFS);
    float* dst = image + tile * (long long)(K * TC_TILE_REFS) + (col >> 3) * 32 + (col & 7) * 4;
    const double* src = y64 + (valid ? j : 0) * d;
    for (int c = 0; c < K / 4; ++c) {
        float w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = 4 * c + i;
            float x = 0.f;
            if (k < d) x = valid ? (float)src[k] : 0.f;
            else if (k == d) x = h_hi;
            else if (k == d + 1) x = h_lo;
            w[i] = x;
        }
        *reinterpret_cast<float4*>(dst + (long long)c * TC_CHUNK_FLOATS) = make_float4(w[0], w[1], w[2], w[3]);
    }
}

}  // namespace

}  // namespace mce
// development: copies the TC_PROBE counters of the current device to out[64] and clears them
extern "C" int mce_tc_probe(unsigned long long* out) {
    if (cudaMemcpyFromSymbol(out, mce::g_tc_probe, 64 * sizeof(unsigned long long)) != cudaSuccess) return -1;
    const unsigned long long zero[64] = {};
    return cudaMemcpyToSymbol(mce::g_tc_probe, zero, sizeof(zero)) == cudaSuccess ? 0 : -1;
}
namespace mce {
// Give-up record of the barrier waits (tcgen05.cuh): copies {flag, a, b, block} to `out` and clears it on the device.
int tc_read_stuck(unsigned int (&out)[4], cudaStream_t st) {
    if (cudaMemcpyFromSymbolAsync(out, g_tc_stuck, sizeof(out), 0, cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
    if (out[0] != 0u) {
        const unsigned int zero[4] = {0u, 0u, 0u, 0u};
        if (cudaMemcpyToSymbol(g_tc_stuck, zero, sizeof(zero)) != cudaSuccess) return -1;
    }
    return 0;
}

size_t tc_image_bytes(long long n_ref, int d) {
    const long long ntiles = (n_ref + TC_TILE_REFS - 1) / TC_TILE_REFS;
    return (size_t)ntiles * tc_kpad(d) * TC_TILE_REFS * sizeof(float);
}

void launch_tc_image(const double* y64, const float* ffma_tiles, long long n_ref, int d, float* image, cudaStream_t st,
                     int* launches) {
    const long long n_pad = (n_ref + TC_TILE_REFS - 1) / TC_TILE_REFS * TC_TILE_REFS;
    if (n_pad == 0) return;
    tc_image_kernel<<<(unsigned)((n_pad + 127) / 128), 128, 0, st>>>(y64, ffma_tiles, n_ref, n_pad, d, tc_kpad(d), image);
    *launches += 1;
}

int tc_filter_query(int d, int list_len, int want_stages, int* queries_per_cta, int* smem_bytes, int* ctas_per_sm, int* stages) {
    if (d < 1 || d > MAX_DIM || list_len > TC_LMAX) return -1;
    const int K = tc_kpad(d);
    const size_t tile = (size_t)K * TC_TILE_REFS * 4;
    const size_t fixed = tc_fixed_smem(K, list_len);
    if (fixed + 2 * tile > 227 * 1024) return -1;
    long long ns = ((long long)TC_SMEM_BUDGET - (long long)fixed) / (long long)tile;
    if (ns > 8) ns = 8;
    if (ns < 2) ns = 2;
    if (want_stages >= 2 && want_stages <= 8 && fixed + (size_t)want_stages * tile <= 227 * 1024) ns = want_stages;
    *queries_per_cta = TC_QUERIES;
    *smem_bytes = (int)(fixed + (size_t)ns * tile);
    *ctas_per_sm = 1;   // a CTA owns all 512 TMEM columns
    *stages = (int)ns;
    return 0;
}

int launch_tc_filter(const TcFilterParams& tp, cudaStream_t st, int* launches) {
    const FilterParams& p = tp.f;
    const int K = tc_kpad(p.d);
    const size_t smem = tc_fixed_smem(K, p.list_len) + (size_t)tp.stages * K * TC_TILE_REFS * 4;
    const unsigned grid = (unsigned)p.nfull + (unsigned)(p.nqblocks - p.nfull) * (unsigned)p.ref_chunks;
    int rc = -1;
    switch (K / 8) {
        case 1: rc = tc_launch_k<1>(tp, grid, smem, st); break;
        case 2: rc = tc_launch_k<2>(tp, grid, smem, st); break;
        case 3: rc = tc_launch_k<3>(tp, grid, smem, st); break;
        case 4: rc = tc_launch_k<4>(tp, grid, smem, st); break;
        case 5: rc = tc_launch_k<5>(tp, grid, smem, st); break;
        case 6: rc = tc_launch_k<6>(tp, grid, smem, st); break;
        case 7: rc = tc_launch_k<7>(tp, grid, smem, st); break;
        case 8: rc = tc_launch_k<8>(tp, grid, smem, st); break;
        case 9: rc = tc_launch_k<9>(tp, grid, smem, st); break;
        default: break;
    }
    if (rc == 0) *launches += 1;
    return rc;
}

}  // namespace mce
