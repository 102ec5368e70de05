// Experiment: can the FP64 pipe of sm_100 (half the FP32 FMA rate on B200) run "evaluation-shaped" instruction mixes
// concurrently with the packed-FP32 scorer mix without slowing it down?  16 warps per block (4 per SM sub-partition);
// the last NF64 warps run the FP64 mix, the others the FFMA2 mix.  Reports evals/s of each kind alone and together.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o corun tools/exp/corun_fp64.cu && ./corun
#include <cuda_runtime.h>
#include <cstdio>

constexpr int kThreads = 512;

__device__ __forceinline__ void mix32(float* out, int iters, float seed, int tid) {      // 2 evals per chain step
    float2 a[4];
    float m[8];
    unsigned cnt = 0;
    float amb = 1e30f;
#pragma unroll
    for (int i = 0; i < 4; ++i) { a[i] = make_float2(seed + i + tid * 1e-9f, seed - i); m[2*i] = seed; m[2*i+1] = seed; }
    const float2 b = make_float2(0.999999f, 0.999998f), c = make_float2(1e-7f * seed, 2e-7f * seed);
    const float g = 1e-30f * seed;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int r = 0; r < 17; ++r) a[i] = __ffma2_rn(a[i], b, c);
            m[2*i]   = fminf(m[2*i],   a[i].x);
            m[2*i+1] = fminf(m[2*i+1], a[i].y);
            cnt += __float_as_uint(a[i].x) >> 31;
            cnt += __float_as_uint(a[i].y) >> 31;
            amb = fminf(amb, fminf(fabsf(a[i].x), fabsf(a[i].y)));
        }
    }
    float s = (float)cnt + (amb <= g ? 1.f : 0.f);
#pragma unroll
    for (int i = 0; i < 4; ++i) s += a[i].x + a[i].y + m[2*i] + m[2*i+1];
    if (s == 123.456f) out[tid] = s;
}

__device__ __forceinline__ void mix64(float* out, int iters, float seed, int tid) {      // 1 eval per chain step
    double a[4], m[4];
    unsigned cnt = 0;
    unsigned amb = 0x7fffffffu;
#pragma unroll
    for (int i = 0; i < 4; ++i) { a[i] = seed + i + tid * 1e-9; m[i] = seed; }
    const double b = 0.999999, c = 1e-7 * seed;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int r = 0; r < 17; ++r) a[i] = __fma_rn(a[i], b, c);
            m[i] = fmin(m[i], a[i]);
            const unsigned hi = (unsigned)__double2hiint(a[i]);
            cnt += hi >> 31;
            amb = min(amb, hi & 0x7fffffffu);
        }
    }
    double s = (double)cnt + (amb <= 12345u ? 1. : 0.);
#pragma unroll
    for (int i = 0; i < 4; ++i) s += a[i] + m[i];
    if (s == 123.456) out[tid] = (float)s;
}

template <int NF64>
__global__ void __launch_bounds__(kThreads, 1) corun(float* out, int it32, int it64, float seed) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int warp = threadIdx.x >> 5;
    if (warp >= 16 - NF64) mix64(out, it64, seed, tid);
    else mix32(out, it32, seed, tid);
}

template <int NF64>
static float run(float* d, int blocks, int it32, int it64) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    corun<NF64><<<blocks, kThreads>>>(d, it32, it64, 1.0f);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        corun<NF64><<<blocks, kThreads>>>(d, it32, it64, 1.0f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

template <int NF64>
static void sweep(float* d, int sms) {
    const int blocks = sms;                       // one 512-thread block per SM
    const int it32 = 4096;
    const double ev32 = (double)blocks * (16 - NF64) * 32 * it32 * 8.0;      // evals by FP32 warps
    if (NF64 == 0) {
        float ms = run<0>(d, blocks, it32, 0);
        printf("{\"nf64\": 0, \"ms\": %.3f, \"gevals32\": %.1f}\n", ms, ev32 / ms * 1e-6);
        return;
    }
    if (NF64 == 16) {
        const int it64 = 2048;
        float ms = run<16>(d, blocks, 0, it64);
        printf("{\"nf64\": 16, \"ms\": %.3f, \"gevals64\": %.1f}\n", ms, (double)blocks * 16 * 32 * it64 * 4.0 / ms * 1e-6);
        return;
    }
    float ms0 = run<NF64>(d, blocks, it32, 0);
    printf("{\"nf64\": %d, \"it64\": 0, \"ms\": %.3f, \"gevals32\": %.1f}\n", NF64, ms0, ev32 / ms0 * 1e-6);
    for (int it64 = 512; it64 <= 16384; it64 *= 2) {
        float ms = run<NF64>(d, blocks, it32, it64);
        const double ev64 = (double)blocks * NF64 * 32 * it64 * 4.0;
        printf("{\"nf64\": %d, \"it64\": %d, \"ms\": %.3f, \"gevals32\": %.1f, \"gevals64\": %.1f, \"total\": %.1f}\n", NF64, it64,
               ms, ev32 / ms * 1e-6, ev64 / ms * 1e-6, (ev32 + ev64) / ms * 1e-6);
    }
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    float* d; cudaMalloc(&d, (size_t)prop.multiProcessorCount * kThreads * 4);
    sweep<0>(d, prop.multiProcessorCount);
    sweep<16>(d, prop.multiProcessorCount);
    sweep<4>(d, prop.multiProcessorCount);
    sweep<8>(d, prop.multiProcessorCount);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
