// General-N least-squares solvers behind the single-call drop-ins lab3.fmatrix_stls(pl, pr) (N >= 8, lab3.py:269-329)
// and pnp.pnp_minimize(X, y, m) (m >= 6, pnp.py:132-152).  Not on the RANSAC hot path (there the sample size is fixed
// at 8 / 6 and f8_solve_* / pnp_solve_jacobi are used); kept on the GPU so the package has no CPU arithmetic at all.
//
// The tall design matrix (N x 9 or 3m x 12) is reduced to its triangular factor by a tree of Givens QR merges
// (each thread folds rows into a private packed upper-triangular R held in registers), then the register-resident
// one-sided Jacobi of jacobi.cuh runs on the small square R; the minimising right singular vector is the same.
#pragma once
#include "pnp_kernels.cuh"

namespace rg {

template <int COLS>
struct PackedR {
    static constexpr int SIZE = COLS * (COLS + 1) / 2;
    __host__ __device__ static constexpr int diag(int i) { return i * COLS - (i * (i - 1)) / 2; }
};

// fold one row into the packed upper-triangular factor with Givens rotations (backward stable)
template <int COLS>
__device__ __forceinline__ void fold_row(double (&R)[PackedR<COLS>::SIZE], double (&r)[COLS]) {
#pragma unroll
    for (int i = 0; i < COLS; ++i) {
        constexpr int dummy = 0; (void)dummy;
        const int d = PackedR<COLS>::diag(i);
        const double a = R[d], b = r[i];
        if (b != 0.0) {
            const double hyp = sqrt(a * a + b * b);
            const double c = a / hyp, s = b / hyp;
            R[d] = hyp;
#pragma unroll
            for (int j = i + 1; j < COLS; ++j) {
                const double t = R[d + (j - i)];
                R[d + (j - i)] = c * t + s * r[j];
                r[j] = c * r[j] - s * t;
            }
        }
    }
}

// Hartley statistics of all N points of both images: out = {a1, b1, c1, a2, b2, c2} (see hartley8)
__global__ void __launch_bounds__(256) stls_stats(const double* __restrict__ pl, const double* __restrict__ pr, int N,
                                                   double* __restrict__ out) {
    __shared__ double red[4][8];
    __shared__ double mean[4];
    double s[4] = {0, 0, 0, 0};
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        s[0] += pl[i]; s[1] += pl[N + i]; s[2] += pr[i]; s[3] += pr[N + i];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = s[k];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
        mean[threadIdx.x] = t / (double)N;
    }
    __syncthreads();
    double q[2] = {0, 0};
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const double dx = pl[i] - mean[0], dy = pl[N + i] - mean[1];
        const double ex = pr[i] - mean[2], ey = pr[N + i] - mean[3];
        q[0] += dx * dx + dy * dy;
        q[1] += ex * ex + ey * ey;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 2; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q[k] += __shfl_xor_sync(0xffffffffu, q[k], o);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = q[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t0 = 0.0, t1 = 0.0;
        for (int w = 0; w < 8; ++w) { t0 += red[0][w]; t1 += red[1][w]; }
        const double L1 = sqrt(t0 / (2.0 * N)), L2 = sqrt(t1 / (2.0 * N));
        out[0] = 1.0 / L1; out[1] = -mean[0] / L1; out[2] = -mean[1] / L1;
        out[3] = 1.0 / L2; out[4] = -mean[2] / L2; out[5] = -mean[3] / L2;
    }
}

struct StlsRows {           // rows of the N x 9 design matrix, lab3.py:312-315
    const double* pl; const double* pr; const double* hp; int N;
    __device__ __forceinline__ void row(int i, double (&o)[9]) const {
        const double Xh = pl[i] * hp[0] + hp[1], Yh = pl[N + i] * hp[0] + hp[2];
        const double xh = pr[i] * hp[3] + hp[4], yh = pr[N + i] * hp[3] + hp[5];
        o[0] = Xh * xh; o[1] = Xh * yh; o[2] = Xh; o[3] = Yh * xh; o[4] = Yh * yh; o[5] = Yh; o[6] = xh; o[7] = yh; o[8] = 1.0;
    }
};

struct PnpRows {            // rows of the 3m x 12 design matrix, pnp.py:138-143
    const double* X; const double* y;
    __device__ __forceinline__ void row(int i, double (&o)[12]) const {
        const int k = i / 3, l = i - 3 * k;
        const double Xh[4] = {X[3 * (size_t)k], X[3 * (size_t)k + 1], X[3 * (size_t)k + 2], 1.0};
        const double y0 = y[2 * (size_t)k], y1 = y[2 * (size_t)k + 1];
        const double r[3] = {l == 0 ? 0.0 : (l == 1 ? 1.0 : -y1), l == 0 ? -1.0 : (l == 1 ? 0.0 : y0),
                             l == 0 ? y1 : (l == 1 ? -y0 : 0.0)};
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) o[4 * a + b] = r[a] * Xh[b];
    }
};

template <int COLS>
struct PackedRows {         // rows of a stack of packed triangular factors
    const double* Rin;
    __device__ __forceinline__ void row(int i, double (&o)[COLS]) const {
        const int t = i / COLS, k = i - t * COLS;
        const double* R = Rin + (size_t)t * PackedR<COLS>::SIZE;
        const int d = PackedR<COLS>::diag(k);
#pragma unroll
        for (int j = 0; j < COLS; ++j) o[j] = (j >= k) ? R[d + (j - k)] : 0.0;
    }
};

// thread t folds rows [t*rows_per_thread, ...) into its own packed R
template <int COLS, class Gen>
__global__ void __launch_bounds__(64) tsqr_stage(Gen gen, int nrows, int rows_per_thread, int nthreads,
                                                  double* __restrict__ Rout) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nthreads) return;
    double R[PackedR<COLS>::SIZE];
#pragma unroll
    for (int k = 0; k < PackedR<COLS>::SIZE; ++k) R[k] = 0.0;
    const int i0 = t * rows_per_thread, i1 = min(i0 + rows_per_thread, nrows);
    for (int i = i0; i < i1; ++i) {
        double r[COLS];
        gen.row(i, r);
        fold_row<COLS>(R, r);
    }
#pragma unroll
    for (int k = 0; k < PackedR<COLS>::SIZE; ++k) Rout[(size_t)t * PackedR<COLS>::SIZE + k] = R[k];
}

// one warp: Jacobi on the final COLS x COLS triangular factor, then the path-specific constraint enforcement
__global__ void __launch_bounds__(32) stls_finish(const double* __restrict__ Rp, const double* __restrict__ hp,
                                                   double* __restrict__ F9) {
    const int lane = threadIdx.x, j = lane & 15, base = lane & 16;
    double w[9], v[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        w[i] = (j < 9 && j >= i) ? Rp[PackedR<9>::diag(i) + (j - i)] : 0.0;     // column j of R
        v[i] = (i == j) ? 1.0 : 0.0;
    }
    GroupJacobi<9, 9>::run(w, v, j, base);
    double s0, s1, smax;
    const int jm = GroupJacobi<9, 9>::smallest(w, j, s0, s1, smax);
    double nv[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) nv[i] = __shfl_sync(0xffffffffu, v[i], base + jm);
    double ss = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) ss += nv[i] * nv[i];
    const double inv = rsqrt(ss);
#pragma unroll
    for (int i = 0; i < 9; ++i) nv[i] *= inv;
    rank2_project(nv);
    double F[9];
    denormalise(nv, hp[0], hp[1], hp[2], hp[3], hp[4], hp[5], F);
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < 9; ++k) F9[k] = F[k];
}

__global__ void __launch_bounds__(32) pnp_min_finish(const double* __restrict__ Rp, double* __restrict__ Rt_out) {
    const int lane = threadIdx.x, j = lane & 15, base = lane & 16;
    double w[12], v[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        w[i] = (j < 12 && j >= i) ? Rp[PackedR<12>::diag(i) + (j - i)] : 0.0;
        v[i] = (i == j) ? 1.0 : 0.0;
    }
    GroupJacobi<12, 12>::run(w, v, j, base);
    double s0, s1, smax;
    const int jm = GroupJacobi<12, 12>::smallest(w, j, s0, s1, smax);
    double c0[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) c0[i] = __shfl_sync(0xffffffffu, v[i], base + jm);
    double Rt[12];
    enforce_pose(c0, Rt);
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < 12; ++k) Rt_out[k] = Rt[k];
}

}  // namespace rg
