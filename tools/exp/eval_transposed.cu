// Experiment (round 2): the epipolar evaluation body with the roles of the two operands swapped.
//   shipped mapping (round 1):  thread <-> 2 hypotheses in registers, correspondences broadcast from shared memory;
//                               the operand shared by consecutive FFMA2 is the 64-bit point pair
//   this file ("transposed"):   thread <-> NP packed correspondence pairs in registers, hypotheses broadcast from shared
//                               memory; the operands shared by consecutive FFMA2 are the 32-bit scalar coefficients
//                               (operand slots A and C), the per-hypothesis count is one REDUX per warp
// Same IEEE op sequence per evaluation as epi_q32 / EpiPolicy::evalN, so the guard-band analysis is unchanged.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I tsbb15-3d-reconstruction-project_b200/csrc \
//        -o tools/exp/eval_transposed tools/exp/eval_transposed.cu
#include "score_core.cuh"
#include <cstdio>
using namespace rg;

constexpr int kHyp = 512;        // hypotheses per shared-memory block (24 KB)

struct HypS { float f[9]; float G; float p0, p1; };     // 48 bytes, same as Hyp32

#ifndef ALU_MODE
#define ALU_MODE 7      // bit0 count, bit1 min|q|, bit2 min(s1,s2) (else s1+s2)
#endif
template <int NP, bool ORDERED>
__device__ __forceinline__ void eval_hyp(const float (&f)[9], const float2 (&x0)[NP], const float2 (&x1)[NP],
                                         const float2 (&y0)[NP], const float2 (&y1)[NP], unsigned& cnt, float& ma, float2& accq) {
    float2 l1x[NP], l1y[NP], l1z[NP], l2x[NP], l2y[NP], r[NP], s1[NP], s2[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) l1x[j] = ffma2_sbs(f[1], y1[j], f[2]);
#pragma unroll
    for (int j = 0; j < NP; ++j) l1y[j] = ffma2_sbs(f[4], y1[j], f[5]);
#pragma unroll
    for (int j = 0; j < NP; ++j) l1z[j] = ffma2_sbs(f[7], y1[j], f[8]);
#pragma unroll
    for (int j = 0; j < NP; ++j) l1x[j] = ffma2_sbc(f[0], y0[j], l1x[j]);
#pragma unroll
    for (int j = 0; j < NP; ++j) l1y[j] = ffma2_sbc(f[3], y0[j], l1y[j]);
#pragma unroll
    for (int j = 0; j < NP; ++j) l1z[j] = ffma2_sbc(f[6], y0[j], l1z[j]);
#pragma unroll
    for (int j = 0; j < NP; ++j) l2x[j] = ffma2_sbs(f[3], x1[j], f[6]);
#pragma unroll
    for (int j = 0; j < NP; ++j) l2y[j] = ffma2_sbs(f[4], x1[j], f[7]);
#pragma unroll
    for (int j = 0; j < NP; ++j) l2x[j] = ffma2_sbc(f[0], x0[j], l2x[j]);
#pragma unroll
    for (int j = 0; j < NP; ++j) l2y[j] = ffma2_sbc(f[1], x0[j], l2y[j]);
#pragma unroll
    for (int j = 0; j < NP; ++j) r[j] = __ffma2_rn(l1y[j], x1[j], l1z[j]);
#pragma unroll
    for (int j = 0; j < NP; ++j) r[j] = __ffma2_rn(l1x[j], x0[j], r[j]);
#pragma unroll
    for (int j = 0; j < NP; ++j) {
        s1[j] = __ffma2_rn(l1x[j], l1x[j], __fmul2_rn(l1y[j], l1y[j]));
        s2[j] = __ffma2_rn(l2x[j], l2x[j], __fmul2_rn(l2y[j], l2y[j]));
    }
#pragma unroll
    for (int j = 0; j < NP; ++j) {
        float2 nm;
        if (ALU_MODE & 4) nm = make_float2(-fminf(s1[j].x, s2[j].x), -fminf(s1[j].y, s2[j].y));
        else { const float2 m = __fadd2_rn(s1[j], s2[j]); nm = make_float2(-m.x, -m.y); }
        const float2 q = __ffma2_rn(r[j], r[j], nm);
        if (ALU_MODE & 1) { cnt += __float_as_uint(q.x) >> 31; cnt += __float_as_uint(q.y) >> 31; }
        if (ALU_MODE & 2) ma = fminf(ma, fminf(fabsf(q.x), fabsf(q.y)));
        if ((ALU_MODE & 3) != 3) { accq = __fadd2_rn(accq, q); }
    }
}

// COUNT: 0 = per-thread counts only (upper bound of the body), 1 = REDUX + lane-0 shared atomic per hypothesis,
//        2 = REDUX + select into the lane that owns hypothesis h % 32 (one shared atomic per lane per 32 hypotheses)
template <int NP, int BLK, int COUNT, int HU, int TH>
__global__ void __launch_bounds__(TH, BLK) bench_t(const float4* __restrict__ pts, const HypS* __restrict__ hyp, int iters,
                                                    int* __restrict__ out, unsigned* __restrict__ list, int* __restrict__ list_n) {
    __shared__ __align__(16) HypS sh[kHyp];
    __shared__ int s_cnt[kHyp];
    for (int i = threadIdx.x; i < kHyp * 3; i += blockDim.x)
        reinterpret_cast<float4*>(sh)[i] = reinterpret_cast<const float4*>(hyp)[i];
    for (int i = threadIdx.x; i < kHyp; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    float2 x0[NP], x1[NP], y0[NP], y1[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) {
        const float4 X = pts[((blockIdx.x * TH + threadIdx.x) * NP + j) * 2];
        const float4 Y = pts[((blockIdx.x * TH + threadIdx.x) * NP + j) * 2 + 1];
        x0[j] = make_float2(X.x, X.y); x1[j] = make_float2(X.z, X.w);
        y0[j] = make_float2(Y.x, Y.y); y1[j] = make_float2(Y.z, Y.w);
    }
    const int lane = threadIdx.x & 31;
    unsigned local = 0;
    float2 accq = make_float2(0.f, 0.f);
    int acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll HU
        for (int h = 0; h < kHyp; ++h) {
            const float4* hp = reinterpret_cast<const float4*>(sh + h);
            const float4 a = hp[0], b = hp[1], c = hp[2];
            const float f[9] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x};
            unsigned cnt = 0;
            float ma = INFINITY;
            eval_hyp<NP, false>(f, x0, x1, y0, y1, cnt, ma, accq);
            if (COUNT == 0) {
                local += cnt;
            } else if (COUNT == 1) {
                const int tot = __reduce_add_sync(0xffffffffu, (int)cnt);
                if (lane == 0) atomicAdd(&s_cnt[h], tot);
            } else {
                const int tot = __reduce_add_sync(0xffffffffu, (int)cnt);
                if (lane == (h & 31)) acc += tot;
                if ((h & 31) == 31) { atomicAdd(&s_cnt[(h & ~31) + lane], acc); acc = 0; }
            }
            if (ma <= c.y) {                                   // guard band hit: rare -> append (hypothesis, thread group)
                const int k = atomicAdd(list_n, 1);
                if (k < 4096) list[k] = ((unsigned)h << 16) | threadIdx.x;
            }
        }
    }
    __syncthreads();
    int s = (int)local + (int)(accq.x + accq.y);
    for (int i = threadIdx.x; i < kHyp; i += blockDim.x) s += s_cnt[i];
    out[blockIdx.x * TH + threadIdx.x] = s;
}

template <int NP, int BLK, int COUNT, int HU, int TH = 256>
static void run(const float4* dp, const HypS* dh, int* dout, unsigned* dl, int* dn, int sms) {
    const int blocks = sms * BLK * 4;
    const int iters = 16;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaMemset(dn, 0, 4);
    bench_t<NP, BLK, COUNT, HU, TH><<<blocks, TH>>>(dp, dh, iters, dout, dl, dn);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        bench_t<NP, BLK, COUNT, HU, TH><<<blocks, TH>>>(dp, dh, iters, dout, dl, dn);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    int n = 0; cudaMemcpy(&n, dn, 4, cudaMemcpyDeviceToHost);
    const double evals = (double)blocks * TH * NP * 2 * kHyp * iters;
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, bench_t<NP, BLK, COUNT, HU, TH>);
    printf("transposed TH %d NP %d blocks/SM %d count-mode %d unroll %d: %.3f ms  %.1f Gevals/s  regs %d  band-hits %d (%s)\n", TH, NP, BLK,
           COUNT, HU, best, evals / best * 1e-6, fa.numRegs, n, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    const size_t npts = (size_t)sms * 4 * 4 * 256 * 8 * 2;           // float4 count upper bound
    float4* hp = new float4[npts];
    for (size_t i = 0; i < npts; ++i)
        hp[i] = make_float4(0.3f + 0.001f * (i % 977), -0.2f + 0.002f * (i % 313), 0.1f * (i % 7), -0.05f * (i % 5));
    HypS* hh = new HypS[kHyp];
    for (int i = 0; i < kHyp; ++i) {
        for (int k = 0; k < 9; ++k) hh[i].f[k] = 0.01f * (float)(((size_t)(i * 12 + k) * 2654435761u) % 201) - 1.0f;
        hh[i].G = 1e-7f; hh[i].p0 = hh[i].p1 = 0.f;
    }
    float4* dp; HypS* dh; int* dout; unsigned* dl; int* dn;
    cudaMalloc(&dp, npts * sizeof(float4)); cudaMalloc(&dh, kHyp * sizeof(HypS));
    cudaMalloc(&dout, (size_t)sms * 16 * 256 * 4); cudaMalloc(&dl, 4096 * 4); cudaMalloc(&dn, 4);
    cudaMemcpy(dp, hp, npts * sizeof(float4), cudaMemcpyHostToDevice);
    cudaMemcpy(dh, hh, kHyp * sizeof(HypS), cudaMemcpyHostToDevice);
    printf("ALU_MODE %d\n", ALU_MODE);
    run<4, 2, 0, 1, 256>(dp, dh, dout, dl, dn, sms);
    run<8, 2, 0, 1, 128>(dp, dh, dout, dl, dn, sms);
    return 0;
}
