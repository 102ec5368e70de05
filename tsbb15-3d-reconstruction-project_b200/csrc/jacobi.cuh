// Register-resident one-sided (Hestenes) Jacobi SVD, one problem per 16-lane group (two problems per warp).
//
// Lane j of the group owns column j of the working matrix W (ROWS doubles, becomes U*Sigma) and column j of V
// (COLS doubles).  A sweep is a round-robin tournament: in each round every lane is paired with one partner, both
// fetch the partner's columns with warp shuffles, compute the same rotation from bit-identical dot products, and each
// updates its own column.  Converges quadratically; the right singular vector of the smallest singular value (the
// null vector for the 8x9 F design matrix, the algebraic minimiser for the 18x12 PnP design matrix) is then the V
// column of the lane whose W column has the smallest norm.
//
// This is the solver BASELINE.json's north_star names.  For the 8-point problem the library's default is the cheaper
// Householder form (f8_solve_qr); both are kept and parity-tested against each other and against LAPACK.
#pragma once
#include "f_kernels.cuh"

namespace rg {

constexpr int kJacobiThreads = 128;     // 8 groups per block
constexpr int kJacobiMaxSweeps = 30;

template <int ROWS, int COLS>
struct GroupJacobi {
    static constexpr int NP = (COLS + 1) & ~1;      // players in the tournament (even)
    static_assert(COLS <= 16, "one column per lane of a 16-lane group");

    // partner of lane j in round r (circle method); lanes >= NP idle
    __device__ static __forceinline__ int partner(int j, int r) {
        if (j >= NP) return j;
        if (j == NP - 1) return r;
        if (j == r) return NP - 1;
        int k = (2 * r - j) % (NP - 1);
        return k < 0 ? k + (NP - 1) : k;
    }

    // returns the number of sweeps executed.  w, v: this lane's columns (zero for lanes >= COLS).
    __device__ static int run(double (&w)[ROWS], double (&v)[COLS], int j, int base) {
        const unsigned full = 0xffffffffu;
        double a = 0.0;
#pragma unroll
        for (int i = 0; i < ROWS; ++i) a += w[i] * w[i];
        double normF2 = a;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) normF2 += __shfl_xor_sync(full, normF2, o);
        const double tiny = 7.9e-31 * normF2;       // (8u)^2 |A|_F^2 : a column this small is rounding noise
        int sweep = 0;
        for (; sweep < kJacobiMaxSweeps; ++sweep) {
            bool rotated = false;
            for (int r = 0; r < NP - 1; ++r) {
                const int pj = partner(j, r);
                const int src = base + pj;
                double wp[ROWS], vp[COLS];
#pragma unroll
                for (int i = 0; i < ROWS; ++i) wp[i] = __shfl_sync(full, w[i], src);
#pragma unroll
                for (int i = 0; i < COLS; ++i) vp[i] = __shfl_sync(full, v[i], src);
                const bool active = (j < COLS) && (pj < COLS) && (pj != j);
                double own = 0.0, oth = 0.0, g = 0.0;
#pragma unroll
                for (int i = 0; i < ROWS; ++i) {
                    own = fma(w[i], w[i], own);
                    oth = fma(wp[i], wp[i], oth);
                    g   = fma(w[i], wp[i], g);
                }
                const bool first = j < pj;
                const double alpha = first ? own : oth;      // squared norm of the lower-indexed column
                const double beta  = first ? oth : own;
                const bool conv = !(fabs(g) > 1e-15 * sqrt(alpha * beta)) || !(fmin(alpha, beta) > tiny);
                if (active && !conv) {
                    double c, s;
                    jacobi_rot(alpha, beta, g, c, s);
                    // lower column' = c*lower - s*upper ; upper column' = s*lower + c*upper
                    const double sg = first ? -s : s;
#pragma unroll
                    for (int i = 0; i < ROWS; ++i) w[i] = fma(sg, wp[i], c * w[i]);
#pragma unroll
                    for (int i = 0; i < COLS; ++i) v[i] = fma(sg, vp[i], c * v[i]);
                    rotated = true;
                }
            }
            if (__ballot_sync(full, rotated) == 0u) { ++sweep; break; }
        }
        return sweep;
    }

    // after run(): index (lane in group) of the column with the smallest norm, plus smallest / second smallest /
    // largest squared norms (i.e. squared singular values) — uniform across the group.
    __device__ static int smallest(const double (&w)[ROWS], int j, double& s_min2, double& s_next2, double& s_max2) {
        const unsigned full = 0xffffffffu;
        double a = 0.0;
#pragma unroll
        for (int i = 0; i < ROWS; ++i) a += w[i] * w[i];
        double key = (j < COLS) ? a : INFINITY;
        if (!(key == key)) key = INFINITY;            // NaN columns never win
        double mn = key;
        int arg = j;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            const double om = __shfl_xor_sync(full, mn, o);
            const int oa = __shfl_xor_sync(full, arg, o);
            if (om < mn || (om == mn && oa < arg)) { mn = om; arg = oa; }
        }
        double nx = (j == arg) ? INFINITY : key;
        double mx = (j < COLS && a == a) ? a : 0.0;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            nx = fmin(nx, __shfl_xor_sync(full, nx, o));
            mx = fmax(mx, __shfl_xor_sync(full, mx, o));
        }
        s_min2 = mn; s_next2 = nx; s_max2 = mx;
        return arg;
    }
};

// 8-point solve, Jacobi form: one hypothesis per 16-lane group (lab3.py:269-329 with np.linalg.svd replaced by the
// one-sided Jacobi above; the null vector is V[:, argmin sigma]).
template <int MODE>
__global__ void __launch_bounds__(kJacobiThreads) f8_solve_jacobi(const double4* __restrict__ pts, const int* __restrict__ idx,
                                                                   unsigned long long seed, unsigned first_pair,
                                                                   const PairInfo* __restrict__ pi,
                                                                   const PairFrame* __restrict__ frames, int P, int Htot,
                                                                   double* __restrict__ F64, Hyp32* __restrict__ hyp32,
                                                                   unsigned char* __restrict__ flags) {
    const int lane = threadIdx.x & 31;
    const int j = lane & 15;
    const int base = lane & 16;
    const int hq = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const bool live = hq < Htot;
    const int h = live ? hq : Htot - 1;
    int lo = 0, hi = P;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (pi[mid].hyp_off <= h) lo = mid; else hi = mid; }
    const PairInfo& info = pi[lo];

    double X[8], Y[8], x[8], y[8];
    bool bad_index = false;
    {
        int id[8];
        if (idx != nullptr) {
            const int4 i0 = reinterpret_cast<const int4*>(idx)[(size_t)h * 2];
            const int4 i1 = reinterpret_cast<const int4*>(idx)[(size_t)h * 2 + 1];
            id[0] = i0.x; id[1] = i0.y; id[2] = i0.z; id[3] = i0.w; id[4] = i1.x; id[5] = i1.y; id[6] = i1.z; id[7] = i1.w;
        } else {      // seeded call: the sample is a pure function of (seed, global pair id, hypothesis index) — philox.cuh
            sample_distinct<8>(seed, first_pair + (unsigned)lo, (unsigned)(info.hyp_first + (h - info.hyp_off)),
                               (unsigned)max(info.n, 8), id);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            int q = id[k];
            bad_index = bad_index || q < 0 || q >= info.n;
            q = q < 0 ? 0 : (q >= info.n ? info.n - 1 : q);
            const double4 v = pts[info.pt_off + q];
            X[k] = v.x; Y[k] = v.y; x[k] = v.z; y[k] = v.w;
        }
    }
    double a1, b1, c1, a2, b2, c2;
    hartley8(X, Y, a1, b1, c1);
    hartley8(x, y, a2, b2, c2);

    double w[8], v[9];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const double Xh = X[k] * a1 + b1, Yh = Y[k] * a1 + c1;
        const double xh = x[k] * a2 + b2, yh = y[k] * a2 + c2;
        const double L = (j < 3) ? Xh : (j < 6 ? Yh : 1.0);
        const int jr = j % 3;
        const double R = (jr == 0) ? xh : (jr == 1 ? yh : 1.0);
        w[k] = (j < 9) ? L * R : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) v[i] = (i == j) ? 1.0 : 0.0;

    GroupJacobi<8, 9>::run(w, v, j, base);
    double s0, s1, smax;
    const int jm = GroupJacobi<8, 9>::smallest(w, j, s0, s1, smax);
    double nv[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) nv[i] = __shfl_sync(0xffffffffu, v[i], base + jm);
    {
        double ss = 0.0;
#pragma unroll
        for (int i = 0; i < 9; ++i) ss += nv[i] * nv[i];
        const double inv = rsqrt(ss);
#pragma unroll
        for (int i = 0; i < 9; ++i) nv[i] *= inv;
    }
    rank2_project(nv);
    double F[9];
    denormalise(nv, a1, b1, c1, a2, b2, c2, F);
    if (live && j == 0) {
        bool finite = true;
#pragma unroll
        for (int k = 0; k < 9; ++k) { F64[(size_t)h * 9 + k] = F[k]; finite = finite && isfinite(F[k]); }
        unsigned char fl = 0;
        if (!(s1 > 1e-18 * smax)) fl |= 1;        // sigma_8 / sigma_1 <= 1e-9: null direction not unique
        if (!finite) fl |= 2;
        if (bad_index) fl |= 4;
        flags[h] = fl;
        make_hyp32<MODE>(F, frames[lo], hyp32 + h);
    }
}

}  // namespace rg
