// Experiment: throughput of the epipolar evaluation body (2 hypotheses per thread, points broadcast from shared memory)
// for different packings of the 17 FP32 lane-ops: which FFMA2 forms are limited by register-file reads on sm_100a?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I tsbb15-3d-reconstruction-project_b200/csrc \
//        -o tools/exp/eval_variants tools/exp/eval_variants.cu
#include "score_core.cuh"
#include <cstdio>
using namespace rg;

struct H2 { float f[9]; };

// VARIANT bit 0: r ops scalar; bit 1: second-step accumulates scalar; bit 2: first-step (s, pt, s) scalar;
//         bit 3: squares/norms scalar; bit 4: final q scalar
template <int VARIANT, int K>
__device__ __forceinline__ void eval_var(const H2 (&H)[K], const float4* __restrict__ pr, unsigned (&cnt)[K], float (&minabs)[K]) {
    const float4 X = pr[0], Y = pr[1];
    const float2 x0 = make_float2(X.x, X.y), x1 = make_float2(X.z, X.w);
    const float2 y0 = make_float2(Y.x, Y.y), y1 = make_float2(Y.z, Y.w);
    float2 l1x[K], l1y[K], l1z[K], l2x[K], l2y[K], r[K], s1[K], s2[K];
    auto sbs = [](float s, float2 b, float c) -> float2 {
        if (VARIANT & 4) return make_float2(__fmaf_rn(s, b.x, c), __fmaf_rn(s, b.y, c));
        return ffma2_sbs(s, b, c);
    };
    auto sbc = [](float s, float2 b, float2 c) -> float2 {
        if (VARIANT & 2) return make_float2(__fmaf_rn(s, b.x, c.x), __fmaf_rn(s, b.y, c.y));
        return ffma2_sbc(s, b, c);
    };
    auto ppp = [](float2 a, float2 b, float2 c) -> float2 {
        if (VARIANT & 1) return make_float2(__fmaf_rn(a.x, b.x, c.x), __fmaf_rn(a.y, b.y, c.y));
        return __ffma2_rn(a, b, c);
    };
#pragma unroll
    for (int k = 0; k < K; ++k) {
        l1x[k] = sbs(H[k].f[1], y1, H[k].f[2]);
        l1y[k] = sbs(H[k].f[4], y1, H[k].f[5]);
        l1z[k] = sbs(H[k].f[7], y1, H[k].f[8]);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        l1x[k] = sbc(H[k].f[0], y0, l1x[k]);
        l1y[k] = sbc(H[k].f[3], y0, l1y[k]);
        l1z[k] = sbc(H[k].f[6], y0, l1z[k]);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        l2x[k] = sbs(H[k].f[3], x1, H[k].f[6]);
        l2y[k] = sbs(H[k].f[4], x1, H[k].f[7]);
        r[k]   = ppp(l1y[k], x1, l1z[k]);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        l2x[k] = sbc(H[k].f[0], x0, l2x[k]);
        l2y[k] = sbc(H[k].f[1], x0, l2y[k]);
        r[k]   = ppp(l1x[k], x0, r[k]);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        if (VARIANT & 8) {
            s1[k] = make_float2(__fmaf_rn(l1x[k].x, l1x[k].x, l1y[k].x * l1y[k].x), __fmaf_rn(l1x[k].y, l1x[k].y, l1y[k].y * l1y[k].y));
            s2[k] = make_float2(__fmaf_rn(l2x[k].x, l2x[k].x, l2y[k].x * l2y[k].x), __fmaf_rn(l2x[k].y, l2x[k].y, l2y[k].y * l2y[k].y));
        } else {
            s1[k] = __ffma2_rn(l1x[k], l1x[k], __fmul2_rn(l1y[k], l1y[k]));
            s2[k] = __ffma2_rn(l2x[k], l2x[k], __fmul2_rn(l2y[k], l2y[k]));
        }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float2 nm = make_float2(-fminf(s1[k].x, s2[k].x), -fminf(s1[k].y, s2[k].y));
        float2 q;
        if (VARIANT & 16) q = make_float2(__fmaf_rn(r[k].x, r[k].x, nm.x), __fmaf_rn(r[k].y, r[k].y, nm.y));
        else q = __ffma2_rn(r[k], r[k], nm);
        cnt[k] += __float_as_uint(q.x) >> 31;
        cnt[k] += __float_as_uint(q.y) >> 31;
        minabs[k] = fminf(minabs[k], fminf(fabsf(q.x), fabsf(q.y)));
    }
}

// ---- ordered variant: every packed op is a volatile asm statement, so ptxas keeps the program order; the order groups
// the instructions that share a point operand (slot B) into runs of 3K so that the operand-reuse cache can serve it ----
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 v_sbs(float s, u64 b, float c) {
    u64 d; asm volatile("{\n.reg .b64 ts, tc;\nmov.b64 ts, {%1, %1};\nmov.b64 tc, {%3, %3};\nfma.rn.f32x2 %0, ts, %2, tc;\n}" : "=l"(d) : "f"(s), "l"(b), "f"(c)); return d; }
__device__ __forceinline__ u64 v_sbc(float s, u64 b, u64 c) {
    u64 d; asm volatile("{\n.reg .b64 ts;\nmov.b64 ts, {%1, %1};\nfma.rn.f32x2 %0, ts, %2, %3;\n}" : "=l"(d) : "f"(s), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 v_ppp(u64 a, u64 b, u64 c) {
    u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 v_mul(u64 a, u64 b) {
    u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// NP point pairs x K hypotheses per call
template <int K, int NP>
__device__ __forceinline__ void eval_ordered(const H2 (&H)[K], const float4* __restrict__ pr, unsigned (&cnt)[K], float (&minabs)[K]) {
    u64 x0[NP], x1[NP], y0[NP], y1[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) {
        const float4 X = pr[2 * j], Y = pr[2 * j + 1];
        x0[j] = pk(X.x, X.y); x1[j] = pk(X.z, X.w); y0[j] = pk(Y.x, Y.y); y1[j] = pk(Y.z, Y.w);
    }
    u64 l1x[NP][K], l1y[NP][K], l1z[NP][K], l2x[NP][K], l2y[NP][K], r[NP][K];
#pragma unroll
    for (int j = 0; j < NP; ++j)
#pragma unroll
        for (int k = 0; k < K; ++k) {
            l1x[j][k] = v_sbs(H[k].f[1], y1[j], H[k].f[2]);
            l1y[j][k] = v_sbs(H[k].f[4], y1[j], H[k].f[5]);
            l1z[j][k] = v_sbs(H[k].f[7], y1[j], H[k].f[8]);
        }
#pragma unroll
    for (int j = 0; j < NP; ++j)
#pragma unroll
        for (int k = 0; k < K; ++k) {
            l1x[j][k] = v_sbc(H[k].f[0], y0[j], l1x[j][k]);
            l1y[j][k] = v_sbc(H[k].f[3], y0[j], l1y[j][k]);
            l1z[j][k] = v_sbc(H[k].f[6], y0[j], l1z[j][k]);
        }
#pragma unroll
    for (int j = 0; j < NP; ++j)
#pragma unroll
        for (int k = 0; k < K; ++k) {
            l2x[j][k] = v_sbs(H[k].f[3], x1[j], H[k].f[6]);
            l2y[j][k] = v_sbs(H[k].f[4], x1[j], H[k].f[7]);
            r[j][k]   = v_ppp(l1y[j][k], x1[j], l1z[j][k]);
        }
#pragma unroll
    for (int j = 0; j < NP; ++j)
#pragma unroll
        for (int k = 0; k < K; ++k) {
            l2x[j][k] = v_sbc(H[k].f[0], x0[j], l2x[j][k]);
            l2y[j][k] = v_sbc(H[k].f[1], x0[j], l2y[j][k]);
            r[j][k]   = v_ppp(l1x[j][k], x0[j], r[j][k]);
        }
#pragma unroll
    for (int j = 0; j < NP; ++j)
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const u64 t1 = v_mul(l1y[j][k], l1y[j][k]);
            const u64 t2 = v_mul(l2y[j][k], l2y[j][k]);
            l1x[j][k] = v_ppp(l1x[j][k], l1x[j][k], t1);      // s1
            l2x[j][k] = v_ppp(l2x[j][k], l2x[j][k], t2);      // s2
        }
#pragma unroll
    for (int j = 0; j < NP; ++j)
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float a0, a1, b0, b1;
            upk(l1x[j][k], a0, a1); upk(l2x[j][k], b0, b1);
            const u64 nm = pk(-fminf(a0, b0), -fminf(a1, b1));
            const u64 q = v_ppp(r[j][k], r[j][k], nm);
            float q0, q1;
            upk(q, q0, q1);
            cnt[k] += __float_as_uint(q0) >> 31;
            cnt[k] += __float_as_uint(q1) >> 31;
            minabs[k] = fminf(minabs[k], fminf(fabsf(q0), fabsf(q1)));
        }
}

// ---- tensor-core upper bound (round 2): the evaluation body WITHOUT the four packed operations of r = x^T F y, r being
// read from shared memory instead (one LDS.64 per two evaluations, standing in for the tcgen05.ld that would fetch the
// tcgen05.mma result from TMEM): the most the FP32 side can gain if the contraction costs nothing at all ----
template <int K>
__device__ __forceinline__ void eval_r_given(const H2 (&H)[K], const float4* __restrict__ pr, const float2* __restrict__ rs,
                                             unsigned (&cnt)[K], float (&minabs)[K]) {
    const float4 X = pr[0], Y = pr[1];
    const float2 x0 = make_float2(X.x, X.y), x1 = make_float2(X.z, X.w);
    const float2 y0 = make_float2(Y.x, Y.y), y1 = make_float2(Y.z, Y.w);
    float2 l1x[K], l1y[K], l2x[K], l2y[K], s1[K], s2[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { l1x[k] = ffma2_sbs(H[k].f[1], y1, H[k].f[2]); l1y[k] = ffma2_sbs(H[k].f[4], y1, H[k].f[5]); }
#pragma unroll
    for (int k = 0; k < K; ++k) { l1x[k] = ffma2_sbc(H[k].f[0], y0, l1x[k]); l1y[k] = ffma2_sbc(H[k].f[3], y0, l1y[k]); }
#pragma unroll
    for (int k = 0; k < K; ++k) { l2x[k] = ffma2_sbs(H[k].f[3], x1, H[k].f[6]); l2y[k] = ffma2_sbs(H[k].f[4], x1, H[k].f[7]); }
#pragma unroll
    for (int k = 0; k < K; ++k) { l2x[k] = ffma2_sbc(H[k].f[0], x0, l2x[k]); l2y[k] = ffma2_sbc(H[k].f[1], x0, l2y[k]); }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        s1[k] = __ffma2_rn(l1x[k], l1x[k], __fmul2_rn(l1y[k], l1y[k]));
        s2[k] = __ffma2_rn(l2x[k], l2x[k], __fmul2_rn(l2y[k], l2y[k]));
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float2 r = rs[k * 256];
        const float2 nm = make_float2(-fminf(s1[k].x, s2[k].x), -fminf(s1[k].y, s2[k].y));
        const float2 q = __ffma2_rn(r, r, nm);
        cnt[k] += __float_as_uint(q.x) >> 31;
        cnt[k] += __float_as_uint(q.y) >> 31;
        minabs[k] = fminf(minabs[k], fminf(fabsf(q.x), fabsf(q.y)));
    }
}

// ---- 16 packed operations per evaluation pair (round 2 experiment): s2 = |l2|^2 is invariant under a per-hypothesis rotation
// of (l2x, l2y), chosen so that l2y' loses its x0 term: l2x' = c x0 + d x1 + e, l2y' = a x1 + b (3 FMA instead of 4); r comes
// from the l1 side as before.  Coefficients: f[0..8] and g[0..4] = (c, d, e, a, b) ----
struct H3 { float f[9]; float g[5]; };
template <int K>
__device__ __forceinline__ void eval_rot(const H3 (&H)[K], const float4* __restrict__ pr, unsigned (&cnt)[K], float (&minabs)[K]) {
    const float4 X = pr[0], Y = pr[1];
    const float2 x0 = make_float2(X.x, X.y), x1 = make_float2(X.z, X.w);
    const float2 y0 = make_float2(Y.x, Y.y), y1 = make_float2(Y.z, Y.w);
    float2 l1x[K], l1y[K], l1z[K], l2x[K], l2y[K], r[K], s1[K], s2[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        l1x[k] = ffma2_sbs(H[k].f[1], y1, H[k].f[2]);
        l1y[k] = ffma2_sbs(H[k].f[4], y1, H[k].f[5]);
        l1z[k] = ffma2_sbs(H[k].f[7], y1, H[k].f[8]);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        l1x[k] = ffma2_sbc(H[k].f[0], y0, l1x[k]);
        l1y[k] = ffma2_sbc(H[k].f[3], y0, l1y[k]);
        l1z[k] = ffma2_sbc(H[k].f[6], y0, l1z[k]);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        l2x[k] = ffma2_sbs(H[k].g[1], x1, H[k].g[2]);
        l2y[k] = ffma2_sbs(H[k].g[3], x1, H[k].g[4]);
        r[k]   = __ffma2_rn(l1y[k], x1, l1z[k]);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        l2x[k] = ffma2_sbc(H[k].g[0], x0, l2x[k]);
        r[k]   = __ffma2_rn(l1x[k], x0, r[k]);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        s1[k] = __ffma2_rn(l1x[k], l1x[k], __fmul2_rn(l1y[k], l1y[k]));
        s2[k] = __ffma2_rn(l2x[k], l2x[k], __fmul2_rn(l2y[k], l2y[k]));
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float2 nm = make_float2(-fminf(s1[k].x, s2[k].x), -fminf(s1[k].y, s2[k].y));
        const float2 q = __ffma2_rn(r[k], r[k], nm);
        cnt[k] += __float_as_uint(q.x) >> 31;
        cnt[k] += __float_as_uint(q.y) >> 31;
        minabs[k] = fminf(minabs[k], fminf(fabsf(q.x), fabsf(q.y)));
    }
}

template <int K, int BLK>
__global__ void __launch_bounds__(256, BLK) bench_rot(const float4* __restrict__ pts, const float* __restrict__ hyp, int iters,
                                                      int* __restrict__ out) {
    __shared__ float4 sp[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) sp[i] = pts[i];
    __syncthreads();
    H3 H[K];
    float G[K];
    unsigned cnt[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float* h = hyp + ((blockIdx.x * 256 + threadIdx.x) * K + k) * 12;
#pragma unroll
        for (int j = 0; j < 9; ++j) H[k].f[j] = h[j];
#pragma unroll
        for (int j = 0; j < 5; ++j) H[k].g[j] = h[(j + 3) % 9] * 0.75f + h[j] * 0.25f;
        G[k] = h[9];
        cnt[k] = 0;
    }
    unsigned flags = 0;
    for (int it = 0; it < iters; ++it) {
        for (int g = 0; g < 64; ++g) {
            float ma[K];
#pragma unroll
            for (int k = 0; k < K; ++k) ma[k] = INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) eval_rot<K>(H, sp + (g * 4 + j) * 2, cnt, ma);
#pragma unroll
            for (int k = 0; k < K; ++k) flags |= (ma[k] <= G[k] ? 1u : 0u) << (g & 31);
        }
    }
    int s = (int)flags;
#pragma unroll
    for (int k = 0; k < K; ++k) s += (int)cnt[k];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}

template <int K, int BLK>
static void run_rot(const float4* dp, const float* dh, int* dout, int sms) {
    const int blocks = sms * 3 * 4;
    const int iters = 64;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench_rot<K, BLK><<<blocks, 256>>>(dp, dh, iters, dout);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        bench_rot<K, BLK><<<blocks, 256>>>(dp, dh, iters, dout);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    const double evals = (double)blocks * 256 * K * 512 * iters;
    printf("16-op body (rotated l2) K %d blocks/SM %d: %.3f ms  %.1f Gevals/s  (%s)\n", K, BLK, best, evals / best * 1e-6, cudaGetErrorString(cudaGetLastError()));
}

constexpr int kPts = 512;      // points per chunk in shared memory (8 KB)

template <int VARIANT, int K>
__global__ void __launch_bounds__(256, 3) bench_kernel(const float4* __restrict__ pts, const float* __restrict__ hyp, int iters,
                                                       int* __restrict__ out) {
    __shared__ float4 sp[kPts];     // kPts/2 point pairs x 2 float4
    __shared__ float2 srs[VARIANT == 200 ? 4 * K * 256 : 1];       // "r from the tensor core": 4 pairs x K hypotheses per thread
    if (VARIANT == 200)
        for (int i = threadIdx.x; i < 4 * K * 256; i += blockDim.x) srs[i] = make_float2(0.25f + 1e-3f * (i % 97), -0.3f + 2e-3f * (i % 89));
    for (int i = threadIdx.x; i < kPts; i += blockDim.x) sp[i] = pts[i];
    __syncthreads();
    H2 H[K];
    float G[K];
    unsigned cnt[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
        for (int j = 0; j < 9; ++j) H[k].f[j] = hyp[((blockIdx.x * 256 + threadIdx.x) * K + k) * 12 + j];
        G[k] = hyp[((blockIdx.x * 256 + threadIdx.x) * K + k) * 12 + 9];
        cnt[k] = 0;
    }
    unsigned flags = 0;
    for (int it = 0; it < iters; ++it) {
        for (int g = 0; g < kPts / 8; ++g) {
            float ma[K];
#pragma unroll
            for (int k = 0; k < K; ++k) ma[k] = INFINITY;
            if (VARIANT == 200) {
#pragma unroll
                for (int j = 0; j < 4; ++j) eval_r_given<K>(H, sp + (g * 4 + j) * 2, srs + j * K * 256 + threadIdx.x, cnt, ma);
            } else if (VARIANT >= 100) {
                constexpr int NP = (VARIANT >= 100 && VARIANT < 200) ? VARIANT - 100 : 1;
#pragma unroll
                for (int j = 0; j < 4; j += NP) eval_ordered<K, NP>(H, sp + (g * 4 + j) * 2, cnt, ma);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) eval_var<VARIANT < 100 ? VARIANT : 0, K>(H, sp + (g * 4 + j) * 2, cnt, ma);
            }
#pragma unroll
            for (int k = 0; k < K; ++k) flags |= (ma[k] <= G[k] ? 1u : 0u) << (g & 31);
        }
    }
    int s = (int)flags;
#pragma unroll
    for (int k = 0; k < K; ++k) s += (int)cnt[k];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}

template <int VARIANT, int K>
static void run(const float4* dp, const float* dh, int* dout, int blocks) {
    const int iters = 64;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench_kernel<VARIANT, K><<<blocks, 256>>>(dp, dh, iters, dout);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        bench_kernel<VARIANT, K><<<blocks, 256>>>(dp, dh, iters, dout);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    const double evals = (double)blocks * 256 * K * kPts * iters;
    printf("variant %2d K %d: %.3f ms  %.1f Gevals/s  (%s)\n", VARIANT, K, best, evals / best * 1e-6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int blocks = prop.multiProcessorCount * 3 * 4;
    float4* hp = new float4[kPts];
    for (int i = 0; i < kPts; ++i) hp[i] = make_float4(0.3f + 0.001f * i, -0.2f + 0.002f * i, 0.1f * (i % 7), -0.05f * (i % 5));
    const size_t nh = (size_t)blocks * 256 * 4 * 12;
    float* hh = new float[nh];
    for (size_t i = 0; i < nh; ++i) hh[i] = 0.01f * (float)((i * 2654435761u) % 201) - 1.0f;
    float4* dp; float* dh; int* dout;
    cudaMalloc(&dp, kPts * sizeof(float4)); cudaMalloc(&dh, nh * 4); cudaMalloc(&dout, (size_t)blocks * 256 * 4);
    cudaMemcpy(dp, hp, kPts * sizeof(float4), cudaMemcpyHostToDevice); cudaMemcpy(dh, hh, nh * 4, cudaMemcpyHostToDevice);
    run_rot<2, 3>(dp, dh, dout, prop.multiProcessorCount);
    run_rot<2, 2>(dp, dh, dout, prop.multiProcessorCount);
    run<200, 2>(dp, dh, dout, blocks);
    run<0, 2>(dp, dh, dout, blocks);
    run<101, 2>(dp, dh, dout, blocks);
    run<102, 2>(dp, dh, dout, blocks);
    run<104, 2>(dp, dh, dout, blocks);
    run<101, 3>(dp, dh, dout, blocks);
    run<102, 3>(dp, dh, dout, blocks);
    run<101, 4>(dp, dh, dout, blocks);
    run<102, 4>(dp, dh, dout, blocks);
    run<0, 2>(dp, dh, dout, blocks);
    run<1, 2>(dp, dh, dout, blocks);
    run<2, 2>(dp, dh, dout, blocks);
    run<3, 2>(dp, dh, dout, blocks);
    run<4, 2>(dp, dh, dout, blocks);
    run<7, 2>(dp, dh, dout, blocks);
    run<8, 2>(dp, dh, dout, blocks);
    run<9, 2>(dp, dh, dout, blocks);
    run<16, 2>(dp, dh, dout, blocks);
    run<17, 2>(dp, dh, dout, blocks);
    run<25, 2>(dp, dh, dout, blocks);
    run<31, 2>(dp, dh, dout, blocks);
    run<0, 1>(dp, dh, dout, blocks);
    run<1, 1>(dp, dh, dout, blocks);
    run<0, 3>(dp, dh, dout, blocks);
    run<1, 3>(dp, dh, dout, blocks);
    return 0;
}
