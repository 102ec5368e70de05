// Two-view geometry either side of the RANSAC hot path, FP64, one problem per thread:
//   * F of a camera pair ........................ lab3.fmatrix_from_cameras  lab3.py:331-351
//   * optimal (Hartley-Sturm) triangulation ..... lab3.triangulate_optimal    lab3.py:382-475
//   * linear triangulation ...................... lab3.triangulate_linear     lab3.py:477-503
//   * relative pose from E + cheirality ......... fun.relative_camera_pose    fun.py:209-258 (fun.specSVD fun.py:186-207)
//   * camera decomposition C = K [R | t] ........ fun.camera_resectioning     fun.py:260-283 (fun.specRQ fun.py:174-184)
//   * first observation within a tolerance ...... tables.Tables.addNewView    tables.py:116-124
// The reference runs these once per correspondence in Python loops (fun.py:352, tables.py:170, 243); here a camera
// pair's correspondences are one launch.  All arithmetic is double precision: the kernels are bound by the FP64 pipe
// (DFMA + MUFU.RCP64H/RSQ64H sequences), not by HBM (32 B in, 24 B out per correspondence against ~10^4 FP64 ops).
#pragma once
#include "pnp_kernels.cuh"

namespace rg {

enum : int { TRI_OPTIMAL = 0, TRI_LINEAR = 1 };

// Per camera pair, produced by geom_prepare.
struct PairGeom {
    double C1[12], C2[12];
    double F[9];             // [C1 n2]_x C1 pinv(C2), |n2| = 1: the reference's F up to its arbitrary SVD sign
    double e1[3], e2[3];     // homogeneous epipoles: e1^T F = 0 (e1 = C1 n2), F e2 = 0 (e2 = C2 n1)
};

// ------------------------------------------------------------------------------------------------
// one-sided Jacobi on a ROWS x COLS matrix held column-wise in registers: W <- W J (orthogonal columns), V <- V J
// ------------------------------------------------------------------------------------------------
template <int ROWS, int COLS>
__device__ __forceinline__ void jacobi_cols(double (&w)[COLS][ROWS], double (&v)[COLS][COLS]) {
    double nf = 0.0;
#pragma unroll
    for (int j = 0; j < COLS; ++j)
#pragma unroll
        for (int i = 0; i < ROWS; ++i) nf = fma(w[j][i], w[j][i], nf);
    const double tiny = 7.9e-31 * nf;                  // (8u)^2 |A|_F^2: a column this small is rounding noise
    for (int sweep = 0; sweep < 30; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int p = 0; p < COLS - 1; ++p)
#pragma unroll
            for (int q = p + 1; q < COLS; ++q) {
                double a = 0.0, b = 0.0, g = 0.0;
#pragma unroll
                for (int i = 0; i < ROWS; ++i) {
                    a = fma(w[p][i], w[p][i], a);
                    b = fma(w[q][i], w[q][i], b);
                    g = fma(w[p][i], w[q][i], g);
                }
                if (fabs(g) > 1e-15 * sqrt(a * b) && fmin(a, b) > tiny) {
                    double c, s;
                    jacobi_rot(a, b, g, c, s);
#pragma unroll
                    for (int i = 0; i < ROWS; ++i) {
                        const double wp = w[p][i], wq = w[q][i];
                        w[p][i] = c * wp - s * wq;
                        w[q][i] = s * wp + c * wq;
                    }
#pragma unroll
                    for (int i = 0; i < COLS; ++i) {
                        const double vp = v[p][i], vq = v[q][i];
                        v[p][i] = c * vp - s * vq;
                        v[q][i] = s * vp + c * vq;
                    }
                    rotated = true;
                }
            }
        if (!rotated) break;
    }
}

__device__ __forceinline__ double det3(double a0, double a1, double a2, double b0, double b1, double b2, double c0,
                                       double c1, double c2) {
    return a0 * (b1 * c2 - b2 * c1) - a1 * (b0 * c2 - b2 * c0) + a2 * (b0 * c1 - b1 * c0);
}

// C: 3x4 row-major, full rank.  n: unit null vector (the camera centre, homogeneous).  pinv: 4x3 row-major (optional).
// SVD of C^T by one-sided Jacobi (same eps * cond accuracy as LAPACK's pinv, no normal equations).
__device__ __forceinline__ void camera_centre_pinv(const double* __restrict__ C, double* __restrict__ n,
                                                   double* __restrict__ pinv) {
    double w[3][4], v[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
#pragma unroll
        for (int i = 0; i < 4; ++i) w[j][i] = C[4 * j + i];
#pragma unroll
        for (int i = 0; i < 3; ++i) v[j][i] = (i == j) ? 1.0 : 0.0;
    }
    jacobi_cols<4, 3>(w, v);
    // normalise the three orthogonal directions, then the 4-D cross product is well scaled
    double inv2[3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
        inv2[j] = 1.0 / (w[j][0] * w[j][0] + w[j][1] * w[j][1] + w[j][2] * w[j][2] + w[j][3] * w[j][3]);
    double q[3][4];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const double s = sqrt(inv2[j]);
#pragma unroll
        for (int i = 0; i < 4; ++i) q[j][i] = w[j][i] * s;
    }
    n[0] = det3(q[0][1], q[0][2], q[0][3], q[1][1], q[1][2], q[1][3], q[2][1], q[2][2], q[2][3]);
    n[1] = -det3(q[0][0], q[0][2], q[0][3], q[1][0], q[1][2], q[1][3], q[2][0], q[2][2], q[2][3]);
    n[2] = det3(q[0][0], q[0][1], q[0][3], q[1][0], q[1][1], q[1][3], q[2][0], q[2][1], q[2][3]);
    n[3] = -det3(q[0][0], q[0][1], q[0][2], q[1][0], q[1][1], q[1][2], q[2][0], q[2][1], q[2][2]);
    const double nn = rsqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2] + n[3] * n[3]);
#pragma unroll
    for (int i = 0; i < 4; ++i) n[i] *= nn;
    if (pinv) {
        // C = sum_j J_j w_j^T  =>  C^+ = sum_j w_j J_j^T / |w_j|^2
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                pinv[3 * r + c] = w[0][r] * v[0][c] * inv2[0] + w[1][r] * v[1][c] * inv2[1] + w[2][r] * v[2][c] * inv2[2];
    }
}

// lab3.fmatrix_from_cameras (lab3.py:331-351) plus both epipoles in homogeneous form.
__device__ __forceinline__ void f_from_cameras(const double* __restrict__ C1, const double* __restrict__ C2,
                                               double* __restrict__ F, double* __restrict__ e1, double* __restrict__ e2) {
    double n2[4], n1[4], pinv[12];
    camera_centre_pinv(C2, n2, pinv);
    camera_centre_pinv(C1, n1, nullptr);
    double A[9];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        e1[i] = C1[4 * i] * n2[0] + C1[4 * i + 1] * n2[1] + C1[4 * i + 2] * n2[2] + C1[4 * i + 3] * n2[3];
        e2[i] = C2[4 * i] * n1[0] + C2[4 * i + 1] * n1[1] + C2[4 * i + 2] * n1[2] + C2[4 * i + 3] * n1[3];
#pragma unroll
        for (int c = 0; c < 3; ++c)
            A[3 * i + c] = C1[4 * i] * pinv[c] + C1[4 * i + 1] * pinv[3 + c] + C1[4 * i + 2] * pinv[6 + c] +
                           C1[4 * i + 3] * pinv[9 + c];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        F[c]     = -e1[2] * A[3 + c] + e1[1] * A[6 + c];
        F[3 + c] =  e1[2] * A[c]     - e1[0] * A[6 + c];
        F[6 + c] = -e1[1] * A[c]     + e1[0] * A[3 + c];
    }
}

// lab3.triangulate_linear (lab3.py:477-503) on homogeneous image points used as they are (not rescaled):
// null direction of the 6x4 matrix [[x1]_x C1; [x2]_x C2] by one-sided Jacobi, X = V[:3, min] / V[3, min].
__device__ __forceinline__ void triangulate_linear_h(const double* __restrict__ C1, const double* __restrict__ C2,
                                                     const double* __restrict__ x1, const double* __restrict__ x2,
                                                     double* __restrict__ X) {
    double w[4][6], v[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        w[j][0] = -x1[2] * C1[4 + j] + x1[1] * C1[8 + j];
        w[j][1] =  x1[2] * C1[j]     - x1[0] * C1[8 + j];
        w[j][2] = -x1[1] * C1[j]     + x1[0] * C1[4 + j];
        w[j][3] = -x2[2] * C2[4 + j] + x2[1] * C2[8 + j];
        w[j][4] =  x2[2] * C2[j]     - x2[0] * C2[8 + j];
        w[j][5] = -x2[1] * C2[j]     + x2[0] * C2[4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[j][i] = (i == j) ? 1.0 : 0.0;
    }
    jacobi_cols<6, 4>(w, v);
    double best = INFINITY;
    double h[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double nj = 0.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) nj = fma(w[j][i], w[j][i], nj);
        const bool take = nj < best || (j == 0);
        if (take) {
            best = nj;
#pragma unroll
            for (int i = 0; i < 4; ++i) h[i] = v[j][i];
        }
    }
    const double inv = 1.0 / h[3];
    X[0] = h[0] * inv; X[1] = h[1] * inv; X[2] = h[2] * inv;
}

// ------------------------------------------------------------------------------------------------
// all complex roots of a real polynomial of degree <= 6 (Aberth-Ehrlich, Bini's Newton-polygon start), real parts only
// ------------------------------------------------------------------------------------------------
struct Cx { double re, im; };
__device__ __forceinline__ Cx cmul(Cx a, Cx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ Cx cdiv(Cx a, Cx b) {
    // Smith's algorithm is not needed: operands are pre-scaled; plain form with one reciprocal
    const double d = 1.0 / (b.re * b.re + b.im * b.im);
    return {(a.re * b.re + a.im * b.im) * d, (a.im * b.re - a.re * b.im) * d};
}
__device__ __forceinline__ Cx cinv(Cx b) {
    const double d = 1.0 / (b.re * b.re + b.im * b.im);
    return {b.re * d, -b.im * d};
}

// a[0..6]: coefficients of t^6 .. t^0 (leading zeros allowed).  Newton correction w = p(z)/p'(z) at z; returns true when
// |p(z)| is below the rounding-error bound of its own evaluation (z is a root to working precision).
__device__ __forceinline__ bool newton_term(const double (&a)[7], Cx z, Cx& w) {
    const double az2 = z.re * z.re + z.im * z.im;
    const double eps = 1.1102230246251565e-16;
    if (az2 <= 1.0) {
        const double az = sqrt(az2);
        Cx p = {a[0], 0.0}, dp = {0.0, 0.0};
        double e = fabs(a[0]);
#pragma unroll
        for (int k = 1; k <= 6; ++k) {
            dp = cmul(dp, z); dp.re += p.re; dp.im += p.im;
            p = cmul(p, z); p.re += a[k];
            e = fma(e, az, fabs(a[k]));
        }
        const double ap2 = p.re * p.re + p.im * p.im;
        const double bound = 28.0 * eps * e;
        if (ap2 <= bound * bound) { w = {0.0, 0.0}; return true; }
        w = cdiv(p, dp);
        return false;
    }
    // |z| > 1: p(z) = z^6 Q(x), x = 1/z, Q(x) = sum a[k] x^k  =>  p/p' = z Q / (6 Q - x Q')
    const Cx x = cinv(z);
    const double ax = rsqrt(az2);
    Cx q = {a[6], 0.0}, dq = {0.0, 0.0};
    double e = fabs(a[6]);
#pragma unroll
    for (int k = 5; k >= 0; --k) {
        dq = cmul(dq, x); dq.re += q.re; dq.im += q.im;
        q = cmul(q, x); q.re += a[k];
        e = fma(e, ax, fabs(a[k]));
    }
    const double aq2 = q.re * q.re + q.im * q.im;
    const double bound = 28.0 * eps * e;
    if (aq2 <= bound * bound) { w = {0.0, 0.0}; return true; }
    const Cx xdq = cmul(x, dq);
    const Cx den = {6.0 * q.re - xdq.re, 6.0 * q.im - xdq.im};
    w = cmul(z, cdiv(q, den));
    return false;
}

// Real parts of the n = 6 - (leading zeros) roots of a (coefficients already stripped of trailing zeros by the caller,
// |a|_max = 1).  The iterates live in shared memory (zs[(2 k + c) * kGeomThreads], c = 0 re / 1 im, one column per
// thread): the root loops stay ROLLED, which keeps the kernel a few thousand instructions long — fully unrolled it was
// 12 400 instructions (198 KB) and spent 11 of every 12 issue slots waiting for the instruction cache (ncu, round 1).
// On return zs[2 k * kGeomThreads], k < n, hold the real parts.  Returns n.
constexpr int kGeomThreads = 128;

__device__ __forceinline__ int sextic_roots_real_parts(const double (&a)[7], double* __restrict__ zs) {
    int lead = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k)
        if (lead == k && a[k] == 0.0) ++lead;
    const int n = 6 - lead;                               // a[6] != 0 here unless the polynomial is identically zero
    if (n <= 0) return 0;
    // ---- starting points: upper convex hull of (i, log2 |c_i|), c_i = coefficient of t^i = a[6 - i], i = 0..n ----
    double lg[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) {                          // starting radii only need ~1e-3 accuracy: float log of the mantissa
        int e;
        const double m = frexp(fabs(a[6 - i]), &e);
        lg[i] = (a[6 - i] != 0.0) ? (double)e + (double)__log2f((float)m) : -1e300;
    }
    {
        int i = 0;                                        // current hull vertex (c_0 = a[6] != 0)
        int seg = 0;
#pragma unroll 1
        while (i < n) {
            double li = 0.0;
#pragma unroll
            for (int t = 0; t < 7; ++t) if (t == i) li = lg[t];
            int jb = i + 1;
            double sb = -INFINITY;
#pragma unroll
            for (int j = 1; j < 7; ++j) {
                if (j > i && j <= n && lg[j] > -1e299) {
                    const double sl = (lg[j] - li) / (double)(j - i);
                    if (sl >= sb) { sb = sl; jb = j; }
                }
            }
            const double fl = floor(-sb);                 // |roots| on this edge ~ (|c_i| / |c_j|)^(1/(j-i)) = 2^-slope
            const double r = ldexp((double)exp2f((float)(-sb - fl)), (int)fmax(fmin(fl, 1000.0), -1000.0));
            const double im = 2.0 / (double)(jb - i);
#pragma unroll 1
            for (int k = i; k < jb; ++k) {
                float sn, cs;
                sincospif((float)((double)(k - i) * im + 0.25 + 0.137 * (double)seg), &sn, &cs);
                zs[(2 * k) * kGeomThreads] = r * cs;
                zs[(2 * k + 1) * kGeomThreads] = r * sn;
            }
            i = jb;
            ++seg;
        }
    }
    // ---- Aberth-Ehrlich iteration (Gauss-Seidel order) ----
    unsigned done = 0u;
    const unsigned all = (1u << n) - 1u;
#pragma unroll 1
    for (int it = 0; it < 60 && done != all; ++it) {
#pragma unroll 1
        for (int k = 0; k < n; ++k) {
            if ((done >> k) & 1u) continue;
            const Cx z = {zs[(2 * k) * kGeomThreads], zs[(2 * k + 1) * kGeomThreads]};
            Cx w;
            if (newton_term(a, z, w)) { done |= 1u << k; continue; }
            Cx sm = {0.0, 0.0};
#pragma unroll 1
            for (int j = 0; j < n; ++j) {
                if (j == k) continue;
                const double dr = z.re - zs[(2 * j) * kGeomThreads], di = z.im - zs[(2 * j + 1) * kGeomThreads];
                const double dd = dr * dr + di * di;
                if (dd > 0.0) {
                    const double id = 1.0 / dd;
                    sm.re += dr * id; sm.im -= di * id;
                }
            }
            const Cx ws = cmul(w, sm);
            const Cx den = {1.0 - ws.re, -ws.im};
            const Cx dz = cdiv(w, den);
            if (isfinite(dz.re) && isfinite(dz.im)) {
                const double zr = z.re - dz.re, zi = z.im - dz.im;
                zs[(2 * k) * kGeomThreads] = zr;
                zs[(2 * k + 1) * kGeomThreads] = zi;
                if (dz.re * dz.re + dz.im * dz.im <= 1e-31 * (zr * zr + zi * zi)) done |= 1u << k;
            } else {
                done |= 1u << k;
            }
        }
    }
    return n;
}

// cost of lab3.py:443-444 with f = f' = 1, evaluated in 1/t for |t| > 1 so that it cannot overflow
__device__ __forceinline__ double hs_cost(double a, double b, double c, double d, double t) {
    double u, v, first;
    if (fabs(t) <= 1.0) {
        u = fma(a, t, b); v = fma(c, t, d);
        first = t * t / (1.0 + t * t);
    } else {
        const double it = 1.0 / t;
        u = fma(b, it, a); v = fma(d, it, c);
        first = 1.0 / (1.0 + it * it);
    }
    const double s = first + v * v / (u * u + v * v);
    return (s == s) ? s : INFINITY;
}

// lab3.triangulate_optimal (lab3.py:382-475) for one correspondence of a prepared camera pair.
__device__ __forceinline__ void triangulate_optimal_pt(const PairGeom& G, double x10, double x11, double x20, double x21,
                                                       double* __restrict__ zs /* shared: 12 x kGeomThreads doubles, + tid */,
                                                       double* __restrict__ X) {
    const double* F = G.F;
    // epipoles of T1^T F T2 are the pair's epipoles moved by -x; the reference divides by the last component and
    // normalises the remaining 2-vector (lab3.py:411-413)
    double c1 = G.e1[0] - x10 * G.e1[2], s1 = G.e1[1] - x11 * G.e1[2];
    double c2 = G.e2[0] - x20 * G.e2[2], s2 = G.e2[1] - x21 * G.e2[2];
    {
        const double n1 = copysign(rsqrt(c1 * c1 + s1 * s1), G.e1[2]);
        const double n2 = copysign(rsqrt(c2 * c2 + s2 * s2), G.e2[2]);
        c1 *= n1; s1 *= n1; c2 *= n2; s2 *= n2;
    }
    // the 2x2 lower-right block of R1 T1^T F T2 R2^T
    const double p0 = F[0] * x20 + F[1] * x21 + F[2];          // (T1^T F T2)[0][2]
    const double p1 = F[3] * x20 + F[4] * x21 + F[5];          //              [1][2]
    const double q0 = F[0] * x10 + F[3] * x11 + F[6];          //              [2][0]
    const double q1 = F[1] * x10 + F[4] * x11 + F[7];          //              [2][1]
    const double d = x10 * p0 + x11 * p1 + (F[6] * x20 + F[7] * x21 + F[8]);
    const double a = -s1 * (-s2 * F[0] + c2 * F[1]) + c1 * (-s2 * F[3] + c2 * F[4]);
    const double b = -s1 * p0 + c1 * p1;
    const double c = -s2 * q0 + c2 * q1;
    // sextic of lab3.py:428-437 (f = f' = 1), scaled to unit max coefficient
    const double k1 = b * c - a * d;
    const double ac2 = a * a + c * c;
    double g[7];
    g[0] = a * c * k1;
    g[1] = ac2 * ac2 + k1 * (b * c + a * d);
    g[2] = 4.0 * ac2 * (a * b + c * d) + 2.0 * a * c * k1 + b * d * k1;
    g[3] = 2.0 * (4.0 * a * b * c * d + 3.0 * a * a * b * b + c * c * (3.0 * d * d + 2.0 * b * b));
    g[4] = -a * a * c * d + a * b * (4.0 * b * b + c * c + 2.0 * d * d) + 2.0 * c * d * (2.0 * d * d + 3.0 * b * b);
    g[5] = b * b * b * b - a * a * d * d + d * d * d * d + b * b * (c * c + 2.0 * d * d);
    g[6] = b * d * k1;
    double gmax = 0.0;
#pragma unroll
    for (int k = 0; k < 7; ++k) gmax = fmax(gmax, fabs(g[k]));
    double best = INFINITY, tb = 0.0;
    bool at_inf = true;
    if (gmax > 0.0 && isfinite(gmax)) {
        const double ig = 1.0 / gmax;
#pragma unroll
        for (int k = 0; k < 7; ++k) g[k] *= ig;
        // np.roots strips exact trailing zeros (roots at t = 0)
        bool zero_root = false;
#pragma unroll
        for (int s = 0; s < 6; ++s) {
            if (g[6] == 0.0) {
#pragma unroll
                for (int k = 6; k > 0; --k) g[k] = g[k - 1];
                g[0] = 0.0;
                zero_root = true;
            }
        }
        const int n = sextic_roots_real_parts(g, zs);
#pragma unroll 1
        for (int k = 0; k < n; ++k) {
            const double tk = zs[(2 * k) * kGeomThreads];
            const double s = hs_cost(a, b, c, d, tk);
            if (s < best) { best = s; tb = tk; at_inf = false; }
        }
        if (zero_root) {
            const double s = hs_cost(a, b, c, d, 0.0);
            if (s < best) { best = s; tb = 0.0; at_inf = false; }
        }
    }
    {
        const double s = 1.0 + c * c / (a * a + c * c);           // asymptote, lab3.py:447
        if (s < best || at_inf) { at_inf = true; }
    }
    double l1[3], l2[3];
    if (!at_inf) {
        l1[0] = -(c * tb + d); l1[1] = a * tb + b; l1[2] = c * tb + d;
        l2[0] = tb; l2[1] = 1.0; l2[2] = -tb;
    } else {
        l1[0] = -c; l1[1] = a; l1[2] = c;
        l2[0] = 1.0; l2[1] = 0.0; l2[2] = -1.0;
    }
    // closest points to the origin on the two lines, moved back: x = T R^T foot(l)   (lab3.py:463-470)
    double x1n[3], x2n[3];
    {
        const double f0 = -l1[0] * l1[2], f1 = -l1[1] * l1[2], f2 = l1[0] * l1[0] + l1[1] * l1[1];
        const double u = c1 * f0 - s1 * f1, v = s1 * f0 + c1 * f1;
        x1n[0] = u + x10 * f2; x1n[1] = v + x11 * f2; x1n[2] = f2;
    }
    {
        const double f0 = -l2[0] * l2[2], f1 = -l2[1] * l2[2], f2 = l2[0] * l2[0] + l2[1] * l2[1];
        const double u = c2 * f0 - s2 * f1, v = s2 * f0 + c2 * f1;
        x2n[0] = u + x20 * f2; x2n[1] = v + x21 * f2; x2n[2] = f2;
    }
    triangulate_linear_h(G.C1, G.C2, x1n, x2n, X);
}

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) geom_prepare(const double* __restrict__ C1, const double* __restrict__ C2, int P,
                                                   PairGeom* __restrict__ out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    PairGeom G;
#pragma unroll
    for (int k = 0; k < 12; ++k) { G.C1[k] = C1[(size_t)p * 12 + k]; G.C2[k] = C2[(size_t)p * 12 + k]; }
    f_from_cameras(G.C1, G.C2, G.F, G.e1, G.e2);
    out[p] = G;
}

__global__ void __launch_bounds__(256) geom_export_F(const PairGeom* __restrict__ G, int P, double* __restrict__ F) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P * 9) F[i] = G[i / 9].F[i % 9];
}

// one correspondence per thread; pair_off (device, P+1) is the CSR table of the camera pairs
template <int METHOD>
__global__ void __launch_bounds__(kGeomThreads) triangulate_kernel(const PairGeom* __restrict__ G, const int* __restrict__ pair_off,
                                                          int P, const double2* __restrict__ x1,
                                                          const double2* __restrict__ x2, int N, double* __restrict__ X) {
    __shared__ PairGeom sG;
    __shared__ double zsh[METHOD == TRI_OPTIMAL ? 12 * kGeomThreads : 1];
    // most launches have one pair per block; the block's first pair is staged in shared memory, others read global
    const int i0 = blockIdx.x * blockDim.x;
    const int pb = find_segment(pair_off, P, min(i0, N - 1));
    for (int k = threadIdx.x; k < (int)(sizeof(PairGeom) / sizeof(double)); k += blockDim.x)
        reinterpret_cast<double*>(&sG)[k] = reinterpret_cast<const double*>(G + pb)[k];
    __syncthreads();
    const int i = i0 + threadIdx.x;
    if (i >= N) return;
    const int p = (i < pair_off[pb + 1]) ? pb : find_segment(pair_off, P, i);
    const PairGeom& g = (p == pb) ? sG : G[p];
    const double2 a = x1[i], b = x2[i];
    double out[3];
    if (METHOD == TRI_OPTIMAL) {
        triangulate_optimal_pt(g, a.x, a.y, b.x, b.y, zsh + threadIdx.x, out);
    } else {
        const double h1[3] = {a.x, a.y, 1.0}, h2[3] = {b.x, b.y, 1.0};
        triangulate_linear_h(g.C1, g.C2, h1, h2, out);
    }
    if (!(isfinite(a.x) && isfinite(a.y) && isfinite(b.x) && isfinite(b.y))) {       // undefined input -> NaN, not Inf
        out[0] = out[1] = out[2] = __longlong_as_double(0x7ff8000000000000ll);
    }
    X[(size_t)i * 3] = out[0]; X[(size_t)i * 3 + 1] = out[1]; X[(size_t)i * 3 + 2] = out[2];
}

// ------------------------------------------------------------------------------------------------
// fun.relative_camera_pose (fun.py:209-258): 4 lanes per pair, one candidate each
// ------------------------------------------------------------------------------------------------
// M: E (P x 9) or, with K9 != null, F with E = K^T F K (fun.py:100-101).  y1, y2: one C-normalised correspondence per
// pair.  Rt (P x 12: R row-major then t) of the first candidate in the reference's order (V W U^T, v3), (V W^T U^T, v3),
// (V W U^T, -v3), (V W^T U^T, -v3) whose optimally triangulated point is in front of both cameras; which[p] = index of
// that candidate, -1 if none (the reference returns None); npass[p] = how many candidates pass.
__global__ void __launch_bounds__(kGeomThreads) relative_pose_kernel(const double* __restrict__ M, const double* __restrict__ K9,
                                                            int k_stride, const double2* __restrict__ y1,
                                                            const double2* __restrict__ y2, int P,
                                                            double* __restrict__ Rt, int* __restrict__ which,
                                                            int* __restrict__ npass) {
    __shared__ double zsh[12 * kGeomThreads];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int pq = tid >> 2;
    const int cand = tid & 3;
    const bool live = pq < P;
    const int p = live ? pq : P - 1;
    double E[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) E[k] = M[(size_t)p * 9 + k];
    if (K9) {
        const double* K = K9 + (size_t)p * k_stride;
        double FK[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) FK[3 * i + j] = E[3 * i] * K[j] + E[3 * i + 1] * K[3 + j] + E[3 * i + 2] * K[6 + j];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) E[3 * i + j] = K[i] * FK[j] + K[3 + i] * FK[3 + j] + K[6 + i] * FK[6 + j];
    }
    // E = U S V^T by one-sided Jacobi: columns of W = E V are sigma_j u_j
    double w[3][3], v[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i) { w[j][i] = E[3 * i + j]; v[j][i] = (i == j) ? 1.0 : 0.0; }
    jacobi3(w, v);
    double sg[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) sg[j] = w[j][0] * w[j][0] + w[j][1] * w[j][1] + w[j][2] * w[j][2];
    int i3 = 0;
    if (sg[1] < sg[i3]) i3 = 1;
    if (sg[2] < sg[i3]) i3 = 2;
    // descending order of the two dominant pairs, as LAPACK returns them
    int ia = (i3 == 0) ? 1 : 0;
    int ib = (i3 == 2) ? 1 : 2;
    {
        double sa = 0.0, sb = 0.0;
#pragma unroll
        for (int j = 0; j < 3; ++j) { if (j == ia) sa = sg[j]; if (j == ib) sb = sg[j]; }
        if (sb > sa) { const int t = ia; ia = ib; ib = t; }
    }
    double u1[3], u2[3], v1[3], v2[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        if (j == ia || j == ib) {
            const double inv = rsqrt(sg[j]);
            double* u = (j == ia) ? u1 : u2;
            double* vv = (j == ia) ? v1 : v2;
#pragma unroll
            for (int i = 0; i < 3; ++i) { u[i] = w[j][i] * inv; vv[i] = v[j][i]; }
        }
    }
    // specSVD (fun.py:186-207): third columns chosen so that det U = det V = +1
    double u3[3], v3[3];
    cross3(u1, u2, u3);
    cross3(v1, v2, v3);
    // V W U^T = -v2 u1^T + v1 u2^T + v3 u3^T ;  V W^T U^T = v2 u1^T - v1 u2^T + v3 u3^T
    const double sw = (cand & 1) ? -1.0 : 1.0;
    const double st = (cand & 2) ? -1.0 : 1.0;
    PairGeom G;
#pragma unroll
    for (int k = 0; k < 12; ++k) G.C1[k] = 0.0;
    G.C1[0] = 1.0; G.C1[5] = 1.0; G.C1[10] = 1.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) G.C2[4 * i + j] = sw * (v1[i] * u2[j] - v2[i] * u1[j]) + v3[i] * u3[j];
        G.C2[4 * i + 3] = st * v3[i];
    }
    f_from_cameras(G.C1, G.C2, G.F, G.e1, G.e2);
    const double2 a = y1[p], b = y2[p];
    double X[3];
    triangulate_optimal_pt(G, a.x, a.y, b.x, b.y, zsh + threadIdx.x, X);
    const double z2 = G.C2[8] * X[0] + G.C2[9] * X[1] + G.C2[10] * X[2] + G.C2[11];
    const bool pass = (X[2] > 0.0) && (z2 > 0.0);
    const unsigned lane = threadIdx.x & 31u;
    const unsigned grp = (__ballot_sync(0xffffffffu, pass) >> (lane & ~3u)) & 0xFu;
    const int first = grp ? (__ffs(grp) - 1) : -1;
    if (live && cand == (first < 0 ? 0 : first)) {
        if (first >= 0) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
#pragma unroll
                for (int j = 0; j < 3; ++j) Rt[(size_t)p * 12 + 3 * i + j] = G.C2[4 * i + j];
                Rt[(size_t)p * 12 + 9 + i] = G.C2[4 * i + 3];
            }
        } else {
            const double qnan = __longlong_as_double(0x7ff8000000000000ll);
#pragma unroll
            for (int k = 0; k < 12; ++k) Rt[(size_t)p * 12 + k] = qnan;
        }
        which[p] = first;
        if (npass) npass[p] = __popc(grp);
    }
}

// ------------------------------------------------------------------------------------------------
// Two-view initialisation of main.py:54-76 for P pairs at once: C-normalise (fun.MakeHomogenous, fun.py:48-55), pick the
// correspondence for the cheirality test, build the cameras [I | 0], [R | t] for the batched triangulation.
// ------------------------------------------------------------------------------------------------
struct Mat3 { double m[9]; };

// pts: (N, 4) pixels (x0, x1, y0, y1).  x1n / x2n: first two components of K^-1 (u, v, 1)^T (NOT divided by the third,
// exactly what main.py passes on: y_hom[:, :2]).  Correspondences with mask == 0 become NaN (they are not triangulated).
__global__ void __launch_bounds__(256) tv_normalise(const double4* __restrict__ pts, const unsigned char* __restrict__ mask,
                                                    int N, Mat3 Kinv, double2* __restrict__ x1n, double2* __restrict__ x2n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double4 v = pts[i];
    double2 a, b;
    a.x = Kinv.m[0] * v.x + Kinv.m[1] * v.y + Kinv.m[2];
    a.y = Kinv.m[3] * v.x + Kinv.m[4] * v.y + Kinv.m[5];
    b.x = Kinv.m[0] * v.z + Kinv.m[1] * v.w + Kinv.m[2];
    b.y = Kinv.m[3] * v.z + Kinv.m[4] * v.w + Kinv.m[5];
    if (mask != nullptr && mask[i] == 0) {
        const double qnan = __longlong_as_double(0x7ff8000000000000ll);
        a = make_double2(qnan, qnan); b = a;
    }
    x1n[i] = a; x2n[i] = b;
}

// one warp per pair: the first correspondence of the pair (main.py:62 uses index 0), or the first with mask != 0
__global__ void __launch_bounds__(128) tv_pick(const double2* __restrict__ x1n, const double2* __restrict__ x2n,
                                               const unsigned char* __restrict__ mask, const int* __restrict__ pair_off,
                                               int P, double2* __restrict__ y1, double2* __restrict__ y2) {
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (p >= P) return;
    const int lo = pair_off[p], hi = pair_off[p + 1];
    int pick = -1;
    if (mask == nullptr) {
        pick = lo < hi ? lo : -1;
    } else {
        for (int base = lo; base < hi && pick < 0; base += 32) {
            const int i = base + lane;
            const unsigned m = __ballot_sync(0xffffffffu, i < hi && mask[i] != 0);
            if (m) pick = base + __ffs(m) - 1;
        }
    }
    if (lane == 0) {
        const double qnan = __longlong_as_double(0x7ff8000000000000ll);
        y1[p] = pick >= 0 ? x1n[pick] : make_double2(qnan, qnan);
        y2[p] = pick >= 0 ? x2n[pick] : make_double2(qnan, qnan);
    }
}

__global__ void __launch_bounds__(256) tv_cameras(const double* __restrict__ Rt, int P, double* __restrict__ C1,
                                                  double* __restrict__ C2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P * 12) return;
    const int p = i / 12, k = i % 12, r = k >> 2, c = k & 3;
    C1[i] = (r == c) ? 1.0 : 0.0;
    C2[i] = (c < 3) ? Rt[(size_t)p * 12 + 3 * r + c] : Rt[(size_t)p * 12 + 9 + r];
}

// ------------------------------------------------------------------------------------------------
// fun.camera_resectioning (fun.py:260-283): C = K [R | t].  The reference's signs come from LAPACK's RQ (dgerqf via
// scipy.linalg.rq): Householder reflectors from the last row upwards, diagonal = -sign(alpha) * norm for rows 2 and 1,
// row 0 untouched.  Restated with the same conventions so that (K, R, t) agree including signs.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) camera_resection_kernel(const double* __restrict__ C, int V, double* __restrict__ K,
                                                              double* __restrict__ R, double* __restrict__ t) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= V) return;
    double A[3][3], bb[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) A[i][j] = C[(size_t)p * 12 + 4 * i + j];
        bb[i] = C[(size_t)p * 12 + 4 * i + 3];
    }
    // Q accumulates the reflectors: A = U Q  =>  Q = H2 H1 applied to I in dorgrq order; here Q is built as the product
    // that was applied to A from the right, transposed: A H2 H1 = U  =>  Q = (H2 H1)^T = H1 H2
    double Hm[3][3] = {{1.0, 0.0, 0.0}, {0.0, 1.0, 0.0}, {0.0, 0.0, 1.0}};       // running product H2 H1 ...
    // ---- reflector for row 2: annihilate A[2][0], A[2][1]; alpha = A[2][2] ----
    {
        const double alpha = A[2][2];
        const double xn = sqrt(A[2][0] * A[2][0] + A[2][1] * A[2][1]);
        if (xn != 0.0) {
            const double beta = -copysign(sqrt(alpha * alpha + xn * xn), alpha);
            const double tau = (beta - alpha) / beta;
            const double sc = 1.0 / (alpha - beta);
            const double vv[3] = {A[2][0] * sc, A[2][1] * sc, 1.0};                  // v, H = I - tau v v^T
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const double dA = tau * (A[i][0] * vv[0] + A[i][1] * vv[1] + A[i][2] * vv[2]);
                const double dH = tau * (Hm[i][0] * vv[0] + Hm[i][1] * vv[1] + Hm[i][2] * vv[2]);
#pragma unroll
                for (int j = 0; j < 3; ++j) { A[i][j] -= dA * vv[j]; Hm[i][j] -= dH * vv[j]; }
            }
            A[2][0] = 0.0; A[2][1] = 0.0; A[2][2] = beta;
        }
    }
    // ---- reflector for row 1: annihilate A[1][0]; alpha = A[1][1]; acts on columns 0, 1 ----
    {
        const double alpha = A[1][1];
        const double xn = fabs(A[1][0]);
        if (xn != 0.0) {
            const double beta = -copysign(sqrt(alpha * alpha + xn * xn), alpha);
            const double tau = (beta - alpha) / beta;
            const double sc = 1.0 / (alpha - beta);
            const double vv[2] = {A[1][0] * sc, 1.0};
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const double dA = tau * (A[i][0] * vv[0] + A[i][1] * vv[1]);
                const double dH = tau * (Hm[i][0] * vv[0] + Hm[i][1] * vv[1]);
                A[i][0] -= dA * vv[0]; A[i][1] -= dA * vv[1];
                Hm[i][0] -= dH * vv[0]; Hm[i][1] -= dH * vv[1];
            }
            A[1][0] = 0.0; A[1][1] = beta;
        }
    }
    // now A = U (upper triangular) and original A = U Hm^T, i.e. Q = Hm^T
    // t = U^-1 b by back substitution (fun.py:266), before U is rescaled
    double tt[3];
    tt[2] = bb[2] / A[2][2];
    tt[1] = (bb[1] - A[1][2] * tt[2]) / A[1][1];
    tt[0] = (bb[0] - A[0][1] * tt[1] - A[0][2] * tt[2]) / A[0][0];
    const double inv = 1.0 / A[2][2];
    double D[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double uii = A[i][i] * inv;
        D[i] = (uii > 0.0) ? 1.0 : (uii < 0.0 ? -1.0 : 0.0);
    }
    const double sgn = (D[0] * D[1] * D[2] == 1.0) ? 1.0 : -1.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            K[(size_t)p * 9 + 3 * i + j] = (j >= i) ? A[i][j] * inv * D[j] : 0.0;
            R[(size_t)p * 9 + 3 * i + j] = sgn * D[i] * Hm[j][i];
        }
        t[(size_t)p * 3 + i] = sgn * D[i] * tt[i];
    }
}

// ------------------------------------------------------------------------------------------------
// tables.py:116-124: for every query row the FIRST observation row within tol (strict <, Euclidean norm); -1 if none.
// grid = (queries / 128, observation segments): each thread scans one segment of the observations (streamed through
// shared memory in tiles) for its query in ascending order and stops at its first hit; the lowest hit over the
// segments wins through an unsigned atomicMin (0xFFFFFFFF = -1 = "none").  Segments that start after an already
// recorded hit are skipped.  Arithmetic as np.linalg.norm evaluates a short vector: s = ((dx*dx) + dy*dy) + dz*dz with
// separate roundings, sqrt(s) < tol — the square root is folded into the threshold: t2 is the smallest double whose
// correctly rounded square root is >= tol (computed on the host), so s < t2 is the SAME predicate, bit for bit.
template <int DIM>
__global__ void __launch_bounds__(128) match_first_kernel(const double* __restrict__ obs, int M, int seg_len,
                                                          const double* __restrict__ y, int N, double t2,
                                                          unsigned* __restrict__ out) {
    constexpr int kTile = 256;
    __shared__ double tile[kTile * DIM];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int m0 = blockIdx.y * seg_len;
    const int m1 = min(M, m0 + seg_len);
    double q[DIM];
#pragma unroll
    for (int k = 0; k < DIM; ++k) q[k] = (i < N) ? y[(size_t)i * DIM + k] : 0.0;
    bool active = i < N;
    if (active && blockIdx.y > 0) active = out[i] >= (unsigned)m0;       // an earlier segment already has a hit
    unsigned hit = 0xFFFFFFFFu;
    for (int base = m0; base < m1; base += kTile) {
        if (__syncthreads_and(!active)) break;
        const int cnt = min(kTile, m1 - base);
        for (int k = threadIdx.x; k < cnt * DIM; k += blockDim.x) tile[k] = obs[(size_t)base * DIM + k];
        __syncthreads();
        if (active) {
            for (int v = 0; v < cnt; ++v) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < DIM; ++k) {
                    const double dlt = __dsub_rn(tile[v * DIM + k], q[k]);
                    s = __dadd_rn(s, __dmul_rn(dlt, dlt));
                }
                if (s < t2) { hit = (unsigned)(base + v); active = false; break; }
            }
        }
    }
    if (hit != 0xFFFFFFFFu) atomicMin(&out[i], hit);
}

}  // namespace rg
