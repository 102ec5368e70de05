// Counter-based random numbers (Philox4x32-10, Salmon et al. SC'11) for the two things the benchmark and the library
// draw on the DEVICE so that they never cross PCIe, yet the host can replay them bit for bit (tsbb15_b200/philox.py is the
// numpy twin of this file; tests compare the two):
//   * sample index sets of the RANSAC loop (reference: fun.py:305-308, `np.random.choice(N, 8, replace=False)` per trial;
//     ransac.py:12-19 for PnP) — k distinct indices of [0, N) per hypothesis, a pure function of
//     (seed, global pair id, hypothesis index);
//   * the synthetic correspondences of BASELINE configs 3-5: X ~ U(Dino bounding box) seen by two
//     cameras, Gaussian pixel noise, the first outlier_frac of the image-2 points replaced by U(image) — a pure function of
//     (seed base + global pair id, point index).
// Everything is integer arithmetic or single IEEE double operations in a fixed order (no FMA contraction: explicit
// __dmul_rn / __dadd_rn), so numpy reproduces it exactly.
#pragma once
#include "common.cuh"

namespace rg {

struct U4 { unsigned x, y, z, w; };

__host__ __device__ __forceinline__ U4 philox4x32_10(U4 ctr, unsigned k0, unsigned k1) {
    const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = (unsigned long long)M0 * ctr.x;
        const unsigned long long p1 = (unsigned long long)M1 * ctr.z;
        U4 n;
        n.x = (unsigned)(p1 >> 32) ^ ctr.y ^ k0;
        n.y = (unsigned)p1;
        n.z = (unsigned)(p0 >> 32) ^ ctr.w ^ k1;
        n.w = (unsigned)p0;
        ctr = n;
        k0 += W0;
        k1 += W1;
    }
    return ctr;
}

// domains (4th counter word) keep the streams of one seed apart
constexpr unsigned kDomSample = 0x53414D50u;   // "SAMP"
constexpr unsigned kDomPoint  = 0x504F494Eu;   // "POIN"
constexpr unsigned kDomCams   = 0x43414D53u;   // "CAMS"

// k distinct indices of [0, n) (n >= k), uniform over ordered k-subsets: draw j takes r = floor(u * (n - j) / 2^32) and maps
// it to the r-th index not chosen before (walk over the sorted earlier choices).  Exactly k 32-bit draws, no rejection
// loop, so the cost does not depend on n and tiny n (n == k) terminates.
template <int K>
__host__ __device__ __forceinline__ void sample_distinct(unsigned long long seed, unsigned pair_id, unsigned h, unsigned n,
                                                         int (&out)[K]) {
    static_assert(K <= 8, "two Philox blocks");
    const unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
    const U4 a = philox4x32_10(U4{h, 0u, pair_id, kDomSample}, k0, k1);
    U4 b = a;
    if (K > 4) b = philox4x32_10(U4{h, 1u, pair_id, kDomSample}, k0, k1);
    const unsigned u[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    unsigned sorted[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        unsigned r = (unsigned)(((unsigned long long)u[j] * (unsigned long long)(n - (unsigned)j)) >> 32);
#pragma unroll
        for (int t = 0; t < j; ++t) r += (r >= sorted[t]) ? 1u : 0u;      // sorted ascending: skip the taken indices
        out[j] = (int)r;
        // insert r into the sorted prefix
        unsigned v = r;
#pragma unroll
        for (int t = 0; t < j; ++t) {
            const unsigned s = sorted[t];
            const bool sw = v < s;
            sorted[t] = sw ? v : s;
            v = sw ? s : v;
        }
        sorted[j] = v;
    }
}

// idx_out (Htot, K) for a batch of pairs described by the pass's PairInfo table: hypothesis h (index inside the pair, plus
// hyp_first in hypothesis-split mode) of the pair with global id first_pair + p
template <int K>
__global__ void __launch_bounds__(256) sample_indices_kernel(const PairInfo* __restrict__ pi, int P, int Htot,
                                                              unsigned long long seed, unsigned first_pair,
                                                              int* __restrict__ idx_out) {
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= Htot) return;
    int lo = 0, hi = P;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (pi[mid].hyp_off <= h) lo = mid; else hi = mid; }
    const PairInfo info = pi[lo];
    int o[K];
    const unsigned n = (unsigned)max(info.n_all, K);
    sample_distinct<K>(seed, first_pair + (unsigned)lo, (unsigned)(info.hyp_first + (h - info.hyp_off)), n, o);
#pragma unroll
    for (int j = 0; j < K; ++j) idx_out[(size_t)h * K + j] = o[j];
}

// ---- synthetic two-view correspondences -----------------------------------------------------------------------
__device__ __forceinline__ double u01(unsigned u) { return __dmul_rn(__dadd_rn((double)u, 0.5), 2.3283064365386963e-10); }  // (u + 0.5) / 2^32
__device__ __forceinline__ double lerp_rn(double lo, double hi, double t) { return __dadd_rn(lo, __dmul_rn(__dadd_rn(hi, -lo), t)); }
// Irwin-Hall(12) from twelve 16-bit uniforms: mean 0, variance 1 - 2^-32 (integer sum: exactly reproducible)
__device__ __forceinline__ double ih12(const unsigned* w6) {
    int s = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) s += (int)(w6[i] & 0xFFFFu) + (int)(w6[i] >> 16);
    return __dmul_rn((double)(s - 393210), 1.52587890625e-05);       // (s - 6 * 65535) / 65536
}
__device__ __forceinline__ void project_rn(const double* __restrict__ C, double X0, double X1, double X2, double& u, double& v) {
    const double a = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(C[0], X0), __dmul_rn(C[1], X1)), __dmul_rn(C[2], X2)), C[3]);
    const double b = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(C[4], X0), __dmul_rn(C[5], X1)), __dmul_rn(C[6], X2)), C[7]);
    const double w = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(C[8], X0), __dmul_rn(C[9], X1)), __dmul_rn(C[10], X2)), C[11]);
    u = __ddiv_rn(a, w);
    v = __ddiv_rn(b, w);
}

struct SynthParams {
    double bbox[6];            // lo/hi of X0, X1, X2
    double sigma_px, width, height;
    int n_out;                 // the first n_out image-2 points of every pair are outliers
    int n_cams;
};

// pair p (global id gp = first_pair + p, seed = seed_base + gp): cameras c1 != c2 drawn from the seed, then N points.
// grid (blocks over points, P)
__global__ void __launch_bounds__(256) synth_two_view_kernel(const double* __restrict__ cams /* n_cams x 12 */, SynthParams prm,
                                                              int N, unsigned long long seed_base, unsigned first_pair,
                                                              double4* __restrict__ pts, int* __restrict__ cam_pair) {
    const unsigned gp = first_pair + blockIdx.y;
    const unsigned long long seed = seed_base + gp;
    const unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
    const U4 cc = philox4x32_10(U4{0u, 0u, gp, kDomCams}, k0, k1);
    const int c1 = (int)(((unsigned long long)cc.x * (unsigned)prm.n_cams) >> 32);
    int c2 = (int)(((unsigned long long)cc.y * (unsigned)(prm.n_cams - 1)) >> 32);
    c2 += (c2 >= c1) ? 1 : 0;
    if (cam_pair != nullptr && blockIdx.x == 0 && threadIdx.x == 0) { cam_pair[2 * blockIdx.y] = c1; cam_pair[2 * blockIdx.y + 1] = c2; }
    const double* C1 = cams + 12 * c1;
    const double* C2 = cams + 12 * c2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        unsigned w[32];
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const U4 r = philox4x32_10(U4{(unsigned)i, (unsigned)b, gp, kDomPoint}, k0, k1);
            w[4 * b] = r.x; w[4 * b + 1] = r.y; w[4 * b + 2] = r.z; w[4 * b + 3] = r.w;
        }
        const double X0 = lerp_rn(prm.bbox[0], prm.bbox[1], u01(w[0]));
        const double X1 = lerp_rn(prm.bbox[2], prm.bbox[3], u01(w[1]));
        const double X2 = lerp_rn(prm.bbox[4], prm.bbox[5], u01(w[2]));
        double xu, xv, yu, yv;
        project_rn(C1, X0, X1, X2, xu, xv);
        project_rn(C2, X0, X1, X2, yu, yv);
        xu = __dadd_rn(xu, __dmul_rn(prm.sigma_px, ih12(w + 8)));
        xv = __dadd_rn(xv, __dmul_rn(prm.sigma_px, ih12(w + 14)));
        yu = __dadd_rn(yu, __dmul_rn(prm.sigma_px, ih12(w + 20)));
        yv = __dadd_rn(yv, __dmul_rn(prm.sigma_px, ih12(w + 26)));
        if (i < prm.n_out) {
            yu = __dmul_rn(prm.width, u01(w[3]));
            yv = __dmul_rn(prm.height, u01(w[4]));
        }
        pts[(size_t)blockIdx.y * N + i] = make_double4(xu, xv, yu, yv);
    }
}

}  // namespace rg
