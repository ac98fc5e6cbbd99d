// FP32 FMA microbenchmark: the measured roofline denominator of the K2 filter.
// MEASURED_PEAKS.json carries HBM and bf16 tensor peaks only; the filter is bound by the
// CUDA-core FMA pipe, so bench.py measures that pipe in the same run.
// variant 0: scalar FFMA, 16 independent accumulator chains per thread
// variant 1: packed FFMA2 (fma.rn.f32x2) with a scalar-broadcast multiplicand, the
//            instruction mix of the filter's inner loop
#include "common.cuh"

namespace mce {

namespace {

constexpr int PEAK_THREADS = 256;
constexpr int PEAK_ACC = 16;
constexpr int PEAK_INNER = 8;

template <int VARIANT>
__global__ void __launch_bounds__(PEAK_THREADS) fma_peak_kernel(float* sink, int iters) {
    const float a0 = 0.999f + 1e-6f * (float)(threadIdx.x & 7);
    const float a1 = 0.998f + 1e-6f * (float)(threadIdx.x & 3);
    if constexpr (VARIANT == 0) {
        float acc[PEAK_ACC];
        float b[4];
#pragma unroll
        for (int i = 0; i < PEAK_ACC; ++i) acc[i] = 1.0f + 0.01f * (float)i;
#pragma unroll
        for (int i = 0; i < 4; ++i) b[i] = 0.001f * (float)(i + 1 + (threadIdx.x & 1));
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < PEAK_INNER; ++u) {
#pragma unroll
                for (int i = 0; i < PEAK_ACC; ++i) acc[i] = fmaf((i & 4) ? a1 : a0, acc[i], b[i & 3]);
            }
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < PEAK_ACC; ++i) s += acc[i];
        if (s == 12345.678f) sink[blockIdx.x * PEAK_THREADS + threadIdx.x] = s;
    } else {
        // VARIANT 1: scalar multiplicand shared by 4 consecutive FFMA2, addend pair cycles
        // VARIANT 2: the filter's pattern -- consecutive pairs share the packed operand, scalars alternate
        // VARIANT 3: one scalar and one packed operand shared by all
        // VARIANT 4: packed x packed (no scalar broadcast)
        float2 acc[PEAK_ACC];
        float2 b[4];
#pragma unroll
        for (int i = 0; i < PEAK_ACC; ++i) acc[i] = make_float2(1.0f + 0.01f * (float)i, 1.0f - 0.01f * (float)i);
#pragma unroll
        for (int i = 0; i < 4; ++i) b[i] = make_float2(0.001f * (float)(i + 1 + (threadIdx.x & 1)), 0.002f);
        const float2 a2 = make_float2(a0, a1);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < PEAK_INNER; ++u) {
#pragma unroll
                for (int i = 0; i < PEAK_ACC; ++i) {
                    if constexpr (VARIANT == 1) {
                        const float a = (i & 4) ? a1 : a0;
                        acc[i] = __ffma2_rn(make_float2(a, a), acc[i], b[i & 3]);
                    } else if constexpr (VARIANT == 2) {
                        const float a = (i & 1) ? a1 : a0;
                        acc[i] = __ffma2_rn(make_float2(a, a), b[(i >> 1) & 3], acc[i]);
                    } else if constexpr (VARIANT == 3) {
                        acc[i] = __ffma2_rn(make_float2(a0, a0), b[0], acc[i]);
                    } else {
                        acc[i] = __ffma2_rn(a2, b[i & 3], acc[i]);
                    }
                }
            }
            if constexpr (VARIANT >= 2) {   // keep values bounded (a*b + acc grows): fold back, once per 128 FFMA2
#pragma unroll
                for (int i = 0; i < PEAK_ACC; ++i) acc[i].x *= 0.5f;
            }
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < PEAK_ACC; ++i) s += acc[i].x + acc[i].y;
        if (s == 12345.678f) sink[blockIdx.x * PEAK_THREADS + threadIdx.x] = s;
    }
}

}  // namespace

// lane-FMAs per thread per iteration
static inline long long fma_per_thread_iter(int variant) { return (long long)PEAK_ACC * PEAK_INNER * (variant ? 2 : 1); }

void launch_fma_peak(int variant, int iters, int grid, float* sink, cudaStream_t st) {
    switch (variant) {
        case 0: fma_peak_kernel<0><<<grid, PEAK_THREADS, 0, st>>>(sink, iters); break;
        case 1: fma_peak_kernel<1><<<grid, PEAK_THREADS, 0, st>>>(sink, iters); break;
        case 2: fma_peak_kernel<2><<<grid, PEAK_THREADS, 0, st>>>(sink, iters); break;
        case 3: fma_peak_kernel<3><<<grid, PEAK_THREADS, 0, st>>>(sink, iters); break;
        default: fma_peak_kernel<4><<<grid, PEAK_THREADS, 0, st>>>(sink, iters); break;
    }
}

long long fma_peak_flops(int variant, int iters, int grid) {
    return 2LL * fma_per_thread_iter(variant) * iters * (long long)grid * PEAK_THREADS;
}

}  // namespace mce
