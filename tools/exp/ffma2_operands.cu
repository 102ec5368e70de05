// Experiment: FFMA2 throughput vs operand pattern (register-file port pressure) on sm_100a.
#include <cuda_runtime.h>
#include <cstdio>
constexpr int T = 256;
template <int MODE>
__global__ void __launch_bounds__(T) k(float* out, int iters, float seed) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    float2 acc[8], X[8], A[8];
    float S[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        acc[i] = make_float2(seed + i, seed - i + tid * 1e-9f);
        X[i] = make_float2(0.999f + i * 1e-4f + tid * 1e-9f, 0.998f - i * 1e-4f);
        A[i] = make_float2(1.0f - i * 1e-5f, 1.0f + i * 1e-5f + tid * 1e-9f);
        S[i] = 1.0f + i * 1e-6f + tid * 1e-9f;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) acc[i] = __ffma2_rn(acc[i], X[0], A[0]);                        // 2 new reads (B,C reused)
                if (MODE == 1) acc[i] = __ffma2_rn(make_float2(S[i], S[i]), X[0], acc[i]);     // scalar + shared pair + acc
                if (MODE == 2) acc[i] = __ffma2_rn(make_float2(S[i], S[i]), X[i], acc[i]);     // scalar + distinct pair + acc
                if (MODE == 3) acc[i] = __ffma2_rn(A[i], X[i], acc[i]);                        // 3 distinct pairs
                if (MODE == 4) acc[i] = __ffma2_rn(A[i], X[0], acc[i]);                        // pair + shared pair + acc
                if (MODE == 5) acc[i] = __ffma2_rn(make_float2(S[i], S[i]), X[i], make_float2(S[(i + 1) & 7], S[(i + 1) & 7]));  // scalar, pair, scalar (no dep)
                if (MODE == 6) acc[i] = __ffma2_rn(acc[i], acc[i], X[i]);                      // square + pair
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
    if (s == 123.456f) out[tid] = s;
}
template <int MODE> void run(float* d, int blocks) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096;
    k<MODE><<<blocks, T>>>(d, iters, 1.f);
    float best = 1e9f;
    for (int r = 0; r < 3; ++r) { cudaEventRecord(e0); k<MODE><<<blocks, T>>>(d, iters, 1.f); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best; }
    double fma = (double)blocks * T * iters * 64.0 * 2.0;
    printf("mode %d: %.1f GFMA/s (%.3f of 36600)\n", MODE, fma / (best * 1e-3) * 1e-9, fma / (best * 1e-3) * 1e-9 / 36600.0);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int blocks = p.multiProcessorCount * 8;
    float* d; cudaMalloc(&d, blocks * T * 4);
    run<0>(d, blocks); run<1>(d, blocks); run<2>(d, blocks); run<3>(d, blocks); run<4>(d, blocks); run<5>(d, blocks); run<6>(d, blocks);
    return 0;
}
