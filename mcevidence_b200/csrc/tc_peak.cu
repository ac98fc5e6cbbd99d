// Unit-peak microbenchmarks for the tensor-core filter engine (knn_filter_tc.cu): the three rates its tile
// loop is bounded by, measured with the instruction shapes the kernel uses.
//   unit 0  FMNMX3      3-input FP32 max on the ALU pipe (the epilogue's threshold test): lane-ops/s
//   unit 1  tcgen05.ld  32x32b.x32 by 8 warps of a CTA owning all 512 TMEM columns: bytes/s into registers
//   unit 2  tcgen05.ld  the same with 4 warps (one per SM sub-partition)
//   unit 3  tcgen05.mma kind::tf32, M = 128, N = 128, K = 8 from shared memory, back to back: FLOP/s
// Not validated on hardware yet (written with the engine, after this round's GPU budget was spent).
#include "common.cuh"
#include "knn_filter.cuh"
#include "tcgen05.cuh"

namespace mce {

namespace {

constexpr int UP_THREADS = 256;
constexpr uint32_t UP_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__global__ void __launch_bounds__(UP_THREADS) fmnmx3_peak_kernel(float* sink, int iters) {
    float m[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) m[i] = (float)((threadIdx.x * 16 + i) % 97) - 40.f;
    const float c0 = -1.0f + (float)(threadIdx.x & 3), c1 = 0.5f * (float)(threadIdx.x & 7);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(m[i]) : "f"(m[(i + 5) & 15]), "f"((u & 1) ? c0 : c1));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += m[i];
    if (s == 12345.678f) sink[blockIdx.x * UP_THREADS + threadIdx.x] = s;
}

// one CTA per SM: TMEM is never written, the loads return whatever it holds
__global__ void __launch_bounds__(UP_THREADS, 1) ldtm_peak_kernel(float* sink, int iters, int warps) {
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    if (warp < warps) {
        uint32_t va[32], vb[32];
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {           // 16 loads of 32 columns: the whole lane quadrant once
                tmem_ld32(base + (uint32_t)(64 * g), va);
                tmem_ld32(base + (uint32_t)(64 * g + 32), vb);
                tmem_wait_ld(va);
                tmem_wait_ld(vb);
                acc ^= va[0] ^ va[31] ^ vb[7] ^ vb[20];
            }
        }
    }
    if (acc == 0x12345678u) sink[blockIdx.x * UP_THREADS + threadIdx.x] = 1.f;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_slot, 512);
    }
}

// one CTA per SM: zero operand tiles in shared memory, four accumulators of 128 columns
__global__ void __launch_bounds__(128, 1) mma_peak_kernel(float* sink, int iters) {
    __shared__ __align__(1024) float a_sm[128 * 8];
    __shared__ __align__(1024) float b_sm[128 * 8];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 128 * 8; i += 128) {
        a_sm[i] = 0.f;
        b_sm[i] = 0.f;
    }
    if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), 512);
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (warp == 0) {
        if (elect_one()) {
            // K-major, no swizzle, one K step: chunk stride 128 rows x 16 B = 2048, row-group stride 128
            const uint64_t adesc = tc_smem_desc(smem_u32(a_sm), 2048, 128);
            const uint64_t bdesc = tc_smem_desc(smem_u32(b_sm), 2048, 128);
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int u = 0; u < 8; ++u) umma_tf32(tmem_base + (uint32_t)((u & 3) * 128), adesc, bdesc, UP_IDESC, 1u);
            }
            umma_commit(smem_u32(&bar));
        }
        __syncwarp();
        tc_wait(smem_u32(&bar), 0);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
    if (iters < 0) sink[0] = 0.f;
}

}  // namespace

// work of one launch in the unit's own measure (lane-ops, bytes, FLOP)
double unit_peak_work(int unit, int iters, int grid) {
    switch (unit) {
        case 0: return (double)iters * 8 * 16 * (double)grid * UP_THREADS;
        case 1: return (double)iters * 16 * 4096.0 * 8 * grid;
        case 2: return (double)iters * 16 * 4096.0 * 4 * grid;
        default: return (double)iters * 8 * 2.0 * 128 * 128 * 8 * grid;
    }
}

int launch_unit_peak(int unit, int iters, int sm_count, float* sink, cudaStream_t st, int* grid_out) {
    int grid = sm_count;
    switch (unit) {
        case 0: grid = sm_count * 8; fmnmx3_peak_kernel<<<grid, UP_THREADS, 0, st>>>(sink, iters); break;
        case 1: ldtm_peak_kernel<<<grid, UP_THREADS, 0, st>>>(sink, iters, 8); break;
        case 2: ldtm_peak_kernel<<<grid, UP_THREADS, 0, st>>>(sink, iters, 4); break;
        case 3: mma_peak_kernel<<<grid, 128, 0, st>>>(sink, iters); break;
        default: return -1;
    }
    *grid_out = grid;
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace mce
