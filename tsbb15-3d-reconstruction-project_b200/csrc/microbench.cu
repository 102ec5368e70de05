// Pipe-throughput micro-benchmarks for sm_100a (B200).
//
// Why this exists: MEASURED_PEAKS.json carries only an HBM copy and a bf16 GEMM figure, but the
// epipolar / reprojection scorers are bound by the FP32 FMA pipes and by warp issue slots
// (SURVEY.md section 8d).  These kernels measure, on the GPU the bench runs on:
//   ffma      : scalar FFMA chains                      -> FP32 peak (2 FLOP / lane / clk)
//   ffma2     : packed fma.rn.f32x2 chains              -> does FFMA2 double the per-issue work?
//   mix       : 18 FFMA  + 3 ALU ops per "eval"         -> scalar scorer issue ceiling
//   mix2      : 9 FFMA2 + 3 ALU ops per "eval"          -> packed scorer ceiling
//   dfma      : FP64 DFMA chains                        -> solver / recheck ceiling
// Built either as a standalone binary (-DRG_MICROBENCH_MAIN) or linked into librg_b200.so where
// rg_measure_peaks() (cabi.cu) calls rg_microbench_run().
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

namespace rgmb {

constexpr int kThreads = 256;
constexpr int kUnroll  = 8;     // independent chains per thread

template <int MODE>
__global__ void __launch_bounds__(kThreads)
pipe_kernel(float* out, int iters, float seed) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (MODE == 0) {                                    // scalar FFMA
        float a[kUnroll];
#pragma unroll
        for (int i = 0; i < kUnroll; ++i) a[i] = seed + i + tid * 1e-9f;
        const float b = 0.999999f, c = 1e-7f * seed;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < kUnroll; ++i) a[i] = __fmaf_rn(a[i], b, c);
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < kUnroll; ++i) s += a[i];
        if (s == 123.456f) out[tid] = s;
    } else if (MODE == 1) {                             // packed FFMA2
        float2 a[kUnroll];
#pragma unroll
        for (int i = 0; i < kUnroll; ++i) a[i] = make_float2(seed + i + tid * 1e-9f, seed - i);
        const float2 b = make_float2(0.999999f, 0.999998f), c = make_float2(1e-7f * seed, 2e-7f * seed);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < kUnroll; ++i) a[i] = __ffma2_rn(a[i], b, c);
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < kUnroll; ++i) s += a[i].x + a[i].y;
        if (s == 123.456f) out[tid] = s;
    } else if (MODE == 2) {                             // 18 FFMA + FMNMX + sign-count + FSETP.OR, 4 chains
        float a[4], m[4];
        unsigned cnt = 0;
        float amb = 1e30f;                               // running min |q| (one FMNMX with |.| modifier)
#pragma unroll
        for (int i = 0; i < 4; ++i) { a[i] = seed + i + tid * 1e-9f; m[i] = seed * 0.5f; }
        const float b = 0.999999f, c = 1e-7f * seed, g = 1e-30f * seed;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int r = 0; r < 18; ++r) a[i] = __fmaf_rn(a[i], b, c);
                m[i] = fminf(m[i], a[i]);
                cnt += __float_as_uint(a[i]) >> 31;
                amb = fminf(amb, fabsf(a[i]));
            }
        }
        float s = (float)cnt + (amb <= g ? 1.f : 0.f);
#pragma unroll
        for (int i = 0; i < 4; ++i) s += a[i] + m[i];
        if (s == 123.456f) out[tid] = s;
    } else if (MODE == 3) {                             // 9 FFMA2 + 3 ALU per eval (2 evals per packed chain)
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
                for (int r = 0; r < 18; ++r) a[i] = __ffma2_rn(a[i], b, c);
                m[2*i]   = fminf(m[2*i],   a[i].x);
                m[2*i+1] = fminf(m[2*i+1], a[i].y);
                cnt += __float_as_uint(a[i].x) >> 31;
                cnt += __float_as_uint(a[i].y) >> 31;
                amb = fminf(amb, fabsf(a[i].x));
                amb = fminf(amb, fabsf(a[i].y));
            }
        }
        float s = (float)cnt + (amb <= g ? 1.f : 0.f);
#pragma unroll
        for (int i = 0; i < 4; ++i) s += a[i].x + a[i].y + m[2*i] + m[2*i+1];
        if (s == 123.456f) out[tid] = s;
    } else {                                            // DFMA
        double a[kUnroll];
#pragma unroll
        for (int i = 0; i < kUnroll; ++i) a[i] = seed + i + tid * 1e-9;
        const double b = 0.999999, c = 1e-7 * seed;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < kUnroll; ++i) a[i] = __fma_rn(a[i], b, c);
        }
        double s = 0.;
#pragma unroll
        for (int i = 0; i < kUnroll; ++i) s += a[i];
        if (s == 123.456) out[tid] = (float)s;
    }
}

struct Result { double ms; double gops; };   // gops: 1e9 "lane operations" per second

template <int MODE>
static Result run_one(float* d_out, int blocks, int iters, double lane_ops_per_thread_iter, cudaStream_t st) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 2; ++w) pipe_kernel<MODE><<<blocks, kThreads, 0, st>>>(d_out, iters, 1.0f);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, st);
        pipe_kernel<MODE><<<blocks, kThreads, 0, st>>>(d_out, iters, 1.0f);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    Result r;
    r.ms = best;
    r.gops = (double)blocks * kThreads * (double)iters * lane_ops_per_thread_iter / (best * 1e-3) * 1e-9;
    return r;
}

}  // namespace rgmb

// out[0] ffma GFMA/s, [1] ffma2 GFMA/s (scalar-FMA equivalents), [2] mix evals/s (G), [3] mix2 evals/s (G),
// [4] dfma GFMA/s, [5] SM count
extern "C" int rg_microbench_run(double* out6, void* stream) {
    using namespace rgmb;
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, dev);
    const int blocks = prop.multiProcessorCount * 8;
    float* d_out = nullptr;
    if (cudaMalloc(&d_out, (size_t)blocks * kThreads * sizeof(float)) != cudaSuccess) return -1;
    const int iters = 4096;
    Result r0 = run_one<0>(d_out, blocks, iters, 8.0 * kUnroll, st);
    Result r1 = run_one<1>(d_out, blocks, iters, 2.0 * 8.0 * kUnroll, st);
    Result r2 = run_one<2>(d_out, blocks, iters / 4, 4.0, st);        // evals
    Result r3 = run_one<3>(d_out, blocks, iters / 4, 8.0, st);        // evals
    Result r4 = run_one<4>(d_out, blocks, iters / 2, 8.0 * kUnroll, st);
    out6[0] = r0.gops; out6[1] = r1.gops; out6[2] = r2.gops; out6[3] = r3.gops; out6[4] = r4.gops;
    out6[5] = prop.multiProcessorCount;
    cudaFree(d_out);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

#ifdef RG_MICROBENCH_MAIN
int main() {
    double o[6];
    int rc = rg_microbench_run(o, nullptr);
    if (rc) { printf("{\"error\": %d}\n", rc); return 1; }
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("{\"sms\": %.0f, \"clock_khz_attr\": %d, \"ffma_gfma_s\": %.1f, \"ffma2_gfma_s\": %.1f, "
           "\"mix_scalar_gevals_s\": %.2f, \"mix_packed_gevals_s\": %.2f, \"dfma_gfma_s\": %.1f, "
           "\"ffma_per_clk_sm_at_attr_clk\": %.2f, \"ffma2_per_clk_sm_at_attr_clk\": %.2f}\n",
           o[5], clk_khz, o[0], o[1], o[2], o[3], o[4],
           o[0] * 1e9 / (o[5] * clk_khz * 1e3), o[1] * 1e9 / (o[5] * clk_khz * 1e3));
    return 0;
}
#endif
