// Host orchestration + C ABI of the two-view geometry kernels (triangulation, relative pose, camera decomposition,
// 2D<->3D observation matching): SURVEY.md section 8f rows N1-N3.
#include "geom_kernels.cuh"
#include "plan.cuh"
#include <cmath>

namespace rg {

// pair table + per-pair geometry live in ctx->geom: [PairGeom x P][int32 pair_off x (P+1)]
static int geom_stage(Ctx* c, cudaStream_t st, int P, const double* C1_dev, const double* C2_dev, const int* pair_off_host,
                      PairGeom** G_out, int** off_out) {
    int rc;
    const size_t gbytes = sizeof(PairGeom) * (size_t)std::max(P, 1);
    if ((rc = ensure(c->geom, gbytes + sizeof(int) * (size_t)(P + 1)))) return rc;
    PairGeom* G = (PairGeom*)c->geom.ptr;
    int* off = (int*)((char*)c->geom.ptr + gbytes);
    if (pair_off_host) {
        RG_CUDA(cudaEventSynchronize(c->staging_free[0]));
        if ((rc = ensure_pinned(c->h_stage[0], sizeof(int) * (size_t)(P + 1)))) return rc;
        memcpy(c->h_stage[0].ptr, pair_off_host, sizeof(int) * (size_t)(P + 1));
        RG_CUDA(cudaMemcpyAsync(off, c->h_stage[0].ptr, sizeof(int) * (size_t)(P + 1), cudaMemcpyHostToDevice, st));
        RG_CUDA(cudaEventRecord(c->staging_free[0], st));
    }
    if (P > 0) {
        geom_prepare<<<ceil_div(P, 64), 64, 0, st>>>(C1_dev, C2_dev, P, G);
        c->last_stats[7] += 1;
        RG_CUDA(cudaGetLastError());
    }
    *G_out = G;
    if (off_out) *off_out = off;
    return RG_OK;
}

static int check_offsets(int P, const int* off) {
    RG_CHECK_ARG(P >= 0, "negative number of camera pairs");
    RG_CHECK_ARG(off != nullptr, "pair_off is null");
    RG_CHECK_ARG(off[0] == 0, "pair_off must start at 0");
    for (int p = 0; p < P; ++p) RG_CHECK_ARG(off[p + 1] >= off[p], "pair_off must be non-decreasing");
    return RG_OK;
}

static int triangulate_dev(Ctx* c, cudaStream_t st, int P, const double* C1, const double* C2, const int* pair_off,
                           const double* x1, const double* x2, int method, double* X) {
    RG_CHECK_ARG(method == TRI_OPTIMAL || method == TRI_LINEAR, "unknown triangulation method");
    int rc = check_offsets(P, pair_off);
    if (rc) return rc;
    RG_CUDA(cudaSetDevice(c->device));
    c->last_stats[7] = 0;
    const int N = P > 0 ? pair_off[P] : 0;
    if (N == 0) return RG_OK;
    RG_CHECK_ARG(C1 && C2 && x1 && x2 && X, "null buffers");
    RG_CHECK_ARG((((uintptr_t)x1 | (uintptr_t)x2) & 15u) == 0, "x1 / x2 must be 16-byte aligned");
    PairGeom* G = nullptr;
    int* off = nullptr;
    if ((rc = geom_stage(c, st, P, C1, C2, pair_off, &G, &off))) return rc;
    const int grid = ceil_div(N, kGeomThreads);
    if (method == TRI_OPTIMAL)
        triangulate_kernel<TRI_OPTIMAL><<<grid, kGeomThreads, 0, st>>>(G, off, P, (const double2*)x1, (const double2*)x2, N, X);
    else
        triangulate_kernel<TRI_LINEAR><<<grid, kGeomThreads, 0, st>>>(G, off, P, (const double2*)x1, (const double2*)x2, N, X);
    c->last_stats[7] += 1;
    RG_CUDA(cudaGetLastError());
    return RG_OK;
}

}  // namespace rg

using namespace rg;

extern "C" {

int rg_triangulate_dev(void* ctx, void* stream, int P, const double* C1_dev, const double* C2_dev,
                       const int* pair_off_host, const double* x1_dev, const double* x2_dev, int method, double* X_dev) {
    RG_CHECK_ARG(ctx != nullptr, "ctx is null");
    return triangulate_dev((Ctx*)ctx, (cudaStream_t)stream, P, C1_dev, C2_dev, pair_off_host, x1_dev, x2_dev, method, X_dev);
}

int rg_triangulate_host(void* ctx, void* stream, int P, const double* C1, const double* C2, const int* pair_off,
                        const double* x1, const double* x2, int method, double* X) {
    RG_CHECK_ARG(ctx != nullptr, "ctx is null");
    Ctx* c = (Ctx*)ctx;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_offsets(P, pair_off);
    if (rc) return rc;
    RG_CUDA(cudaSetDevice(c->device));
    const size_t N = P > 0 ? (size_t)pair_off[P] : 0;
    if (N == 0) return RG_OK;
    RG_CHECK_ARG(C1 && C2 && x1 && x2 && X, "null buffers");
    if ((rc = ensure(c->d_in_a, sizeof(double) * 4 * N))) return rc;
    if ((rc = ensure(c->d_in_b, sizeof(double) * 24 * (size_t)P))) return rc;
    if ((rc = ensure(c->d_out_b, sizeof(double) * 3 * N))) return rc;
    double* dx1 = (double*)c->d_in_a.ptr;
    double* dx2 = dx1 + 2 * N;
    double* dC1 = (double*)c->d_in_b.ptr;
    double* dC2 = dC1 + 12 * (size_t)P;
    RG_CUDA(cudaMemcpyAsync(dx1, x1, sizeof(double) * 2 * N, cudaMemcpyHostToDevice, st));
    RG_CUDA(cudaMemcpyAsync(dx2, x2, sizeof(double) * 2 * N, cudaMemcpyHostToDevice, st));
    RG_CUDA(cudaMemcpyAsync(dC1, C1, sizeof(double) * 12 * (size_t)P, cudaMemcpyHostToDevice, st));
    RG_CUDA(cudaMemcpyAsync(dC2, C2, sizeof(double) * 12 * (size_t)P, cudaMemcpyHostToDevice, st));
    if ((rc = triangulate_dev(c, st, P, dC1, dC2, pair_off, dx1, dx2, method, (double*)c->d_out_b.ptr))) return rc;
    RG_CUDA(cudaMemcpyAsync(X, c->d_out_b.ptr, sizeof(double) * 3 * N, cudaMemcpyDeviceToHost, st));
    RG_CUDA(cudaStreamSynchronize(st));
    return RG_OK;
}

int rg_fmatrix_from_cameras_host(void* ctx, void* stream, int P, const double* C1, const double* C2, double* F) {
    RG_CHECK_ARG(ctx != nullptr, "ctx is null");
    RG_CHECK_ARG(P >= 0, "negative number of camera pairs");
    Ctx* c = (Ctx*)ctx;
    cudaStream_t st = (cudaStream_t)stream;
    RG_CUDA(cudaSetDevice(c->device));
    if (P == 0) return RG_OK;
    RG_CHECK_ARG(C1 && C2 && F, "null buffers");
    int rc;
    if ((rc = ensure(c->d_in_b, sizeof(double) * 24 * (size_t)P))) return rc;
    if ((rc = ensure(c->d_out_b, sizeof(double) * 9 * (size_t)P))) return rc;
    double* dC1 = (double*)c->d_in_b.ptr;
    double* dC2 = dC1 + 12 * (size_t)P;
    RG_CUDA(cudaMemcpyAsync(dC1, C1, sizeof(double) * 12 * (size_t)P, cudaMemcpyHostToDevice, st));
    RG_CUDA(cudaMemcpyAsync(dC2, C2, sizeof(double) * 12 * (size_t)P, cudaMemcpyHostToDevice, st));
    c->last_stats[7] = 0;
    PairGeom* G = nullptr;
    if ((rc = geom_stage(c, st, P, dC1, dC2, nullptr, &G, nullptr))) return rc;
    geom_export_F<<<ceil_div((long long)P * 9, 256), 256, 0, st>>>(G, P, (double*)c->d_out_b.ptr);
    c->last_stats[7] += 1;
    RG_CUDA(cudaGetLastError());
    RG_CUDA(cudaMemcpyAsync(F, c->d_out_b.ptr, sizeof(double) * 9 * (size_t)P, cudaMemcpyDeviceToHost, st));
    RG_CUDA(cudaStreamSynchronize(st));
    return RG_OK;
}

int rg_relative_pose_dev(void* ctx, void* stream, int P, const double* M_dev, const double* K_dev, int k_per_pair,
                         const double* y1_dev, const double* y2_dev, double* Rt_dev, int32_t* which_dev,
                         int32_t* npass_dev) {
    RG_CHECK_ARG(ctx != nullptr, "ctx is null");
    RG_CHECK_ARG(P >= 0, "negative number of pairs");
    Ctx* c = (Ctx*)ctx;
    cudaStream_t st = (cudaStream_t)stream;
    RG_CUDA(cudaSetDevice(c->device));
    c->last_stats[7] = 0;
    if (P == 0) return RG_OK;
    RG_CHECK_ARG(M_dev && y1_dev && y2_dev && Rt_dev && which_dev, "null buffers");
    RG_CHECK_ARG((((uintptr_t)y1_dev | (uintptr_t)y2_dev) & 15u) == 0, "y1 / y2 must be 16-byte aligned");
    relative_pose_kernel<<<ceil_div((long long)P * 4, kGeomThreads), kGeomThreads, 0, st>>>(M_dev, K_dev, k_per_pair ? 9 : 0,
                                                                          (const double2*)y1_dev, (const double2*)y2_dev, P,
                                                                          Rt_dev, which_dev, npass_dev);
    c->last_stats[7] += 1;
    RG_CUDA(cudaGetLastError());
    return RG_OK;
}

int rg_relative_pose_host(void* ctx, void* stream, int P, const double* M, const double* K, int k_per_pair,
                          const double* y1, const double* y2, double* Rt, int32_t* which, int32_t* npass) {
    RG_CHECK_ARG(ctx != nullptr, "ctx is null");
    RG_CHECK_ARG(P >= 0, "negative number of pairs");
    Ctx* c = (Ctx*)ctx;
    cudaStream_t st = (cudaStream_t)stream;
    RG_CUDA(cudaSetDevice(c->device));
    if (P == 0) return RG_OK;
    RG_CHECK_ARG(M && y1 && y2 && Rt && which, "null buffers");
    const size_t p = (size_t)P;
    const size_t nk = K ? (k_per_pair ? 9 * p : 9) : 0;
    int rc;
    if ((rc = ensure(c->d_in_a, sizeof(double) * (4 * p + 9 * p + nk)))) return rc;
    if ((rc = ensure(c->d_out_b, sizeof(double) * 12 * p + sizeof(int) * 2 * p))) return rc;
    double* dy1 = (double*)c->d_in_a.ptr;          // double2 loads: the 16-byte aligned arrays come first
    double* dy2 = dy1 + 2 * p;
    double* dM = dy2 + 2 * p;
    double* dK = dM + 9 * p;
    double* dRt = (double*)c->d_out_b.ptr;
    int* dwhich = (int*)(dRt + 12 * p);
    int* dnpass = dwhich + p;
    RG_CUDA(cudaMemcpyAsync(dM, M, sizeof(double) * 9 * p, cudaMemcpyHostToDevice, st));
    if (K) RG_CUDA(cudaMemcpyAsync(dK, K, sizeof(double) * nk, cudaMemcpyHostToDevice, st));
    RG_CUDA(cudaMemcpyAsync(dy1, y1, sizeof(double) * 2 * p, cudaMemcpyHostToDevice, st));
    RG_CUDA(cudaMemcpyAsync(dy2, y2, sizeof(double) * 2 * p, cudaMemcpyHostToDevice, st));
    if ((rc = rg_relative_pose_dev(ctx, stream, P, dM, K ? dK : nullptr, k_per_pair, dy1, dy2, dRt, dwhich, dnpass)))
        return rc;
    RG_CUDA(cudaMemcpyAsync(Rt, dRt, sizeof(double) * 12 * p, cudaMemcpyDeviceToHost, st));
    RG_CUDA(cudaMemcpyAsync(which, dwhich, sizeof(int) * p, cudaMemcpyDeviceToHost, st));
    if (npass) RG_CUDA(cudaMemcpyAsync(npass, dnpass, sizeof(int) * p, cudaMemcpyDeviceToHost, st));
    RG_CUDA(cudaStreamSynchronize(st));
    return RG_OK;
}

// main.py:54-76 for P pairs: E = K^T F K, C-normalised points, relative pose (cheirality on the pair's first
// correspondence, or first inlier when a mask is given), optimal triangulation of every (inlier) correspondence
int rg_two_view_init_dev(void* ctx, void* stream, int P, const double* pts64_dev, const int32_t* pair_off_host,
                         const double* F_dev, const double* K9_host, const unsigned char* mask_dev, double* Rt_dev,
                         int32_t* which_dev, double* X_dev) {
    RG_CHECK_ARG(ctx != nullptr, "ctx is null");
    Ctx* c = (Ctx*)ctx;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_offsets(P, pair_off_host);
    if (rc) return rc;
    RG_CHECK_ARG(K9_host != nullptr, "K is null");
    RG_CUDA(cudaSetDevice(c->device));
    c->last_stats[7] = 0;
    if (P == 0) return RG_OK;
    const size_t N = (size_t)pair_off_host[P], p = (size_t)P;
    RG_CHECK_ARG(F_dev && Rt_dev && which_dev && (N == 0 || (pts64_dev && X_dev)), "null buffers");
    // K^-1 by the adjugate (3x3, host)
    const double* K = K9_host;
    Mat3 Ki;
    {
        const double c00 = K[4] * K[8] - K[5] * K[7], c01 = K[5] * K[6] - K[3] * K[8], c02 = K[3] * K[7] - K[4] * K[6];
        const double det = K[0] * c00 + K[1] * c01 + K[2] * c02;
        RG_CHECK_ARG(det != 0.0 && std::isfinite(det), "K is singular");
        const double id = 1.0 / det;
        Ki.m[0] = c00 * id; Ki.m[1] = (K[2] * K[7] - K[1] * K[8]) * id; Ki.m[2] = (K[1] * K[5] - K[2] * K[4]) * id;
        Ki.m[3] = c01 * id; Ki.m[4] = (K[0] * K[8] - K[2] * K[6]) * id; Ki.m[5] = (K[2] * K[3] - K[0] * K[5]) * id;
        Ki.m[6] = c02 * id; Ki.m[7] = (K[1] * K[6] - K[0] * K[7]) * id; Ki.m[8] = (K[0] * K[4] - K[1] * K[3]) * id;
    }
    // workspace: x1n, x2n (N double2 each) | y1, y2 (P double2 each) | C1, C2 (P x 12) | K (9) | pair_off (P + 1 ints)
    const size_t nd = 4 * N + 4 * p + 24 * p + 10;
    if ((rc = ensure(c->geom_ws, sizeof(double) * nd + sizeof(int) * (p + 1)))) return rc;
    double* w = (double*)c->geom_ws.ptr;
    double2* x1n = (double2*)w;
    double2* x2n = x1n + N;
    double2* y1 = x2n + N;
    double2* y2 = y1 + p;
    double* C1 = (double*)(y2 + p);
    double* C2 = C1 + 12 * p;
    double* dK = C2 + 12 * p;
    int* doff = (int*)(dK + 10);
    RG_CUDA(cudaEventSynchronize(c->staging_free[0]));
    if ((rc = ensure_pinned(c->h_stage[0], sizeof(double) * 10 + sizeof(int) * (p + 1)))) return rc;
    memcpy(c->h_stage[0].ptr, K, sizeof(double) * 9);
    memcpy((char*)c->h_stage[0].ptr + sizeof(double) * 10, pair_off_host, sizeof(int) * (p + 1));
    RG_CUDA(cudaMemcpyAsync(dK, c->h_stage[0].ptr, sizeof(double) * 10 + sizeof(int) * (p + 1), cudaMemcpyHostToDevice, st));
    RG_CUDA(cudaEventRecord(c->staging_free[0], st));
    if (N) tv_normalise<<<ceil_div((long long)N, 256), 256, 0, st>>>((const double4*)pts64_dev, mask_dev, (int)N, Ki, x1n, x2n);
    tv_pick<<<ceil_div((long long)P * 32, 128), 128, 0, st>>>(x1n, x2n, mask_dev, doff, P, y1, y2);
    relative_pose_kernel<<<ceil_div((long long)P * 4, kGeomThreads), kGeomThreads, 0, st>>>(F_dev, dK, 0, y1, y2, P, Rt_dev,
                                                                                            which_dev, nullptr);
    tv_cameras<<<ceil_div((long long)P * 12, 256), 256, 0, st>>>(Rt_dev, P, C1, C2);
    RG_CUDA(cudaGetLastError());
    const int launches = 4;
    if (N) {
        if ((rc = triangulate_dev(c, st, P, C1, C2, pair_off_host, (const double*)x1n, (const double*)x2n, TRI_OPTIMAL, X_dev)))
            return rc;
    }
    c->last_stats[7] += launches;
    return RG_OK;
}

int rg_two_view_init_host(void* ctx, void* stream, int P, const double* pts64, const int32_t* pair_off, const double* F,
                          const double* K9, const unsigned char* mask, double* Rt, int32_t* which, double* X) {
    RG_CHECK_ARG(ctx != nullptr, "ctx is null");
    Ctx* c = (Ctx*)ctx;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_offsets(P, pair_off);
    if (rc) return rc;
    RG_CUDA(cudaSetDevice(c->device));
    if (P == 0) return RG_OK;
    const size_t N = (size_t)pair_off[P], p = (size_t)P;
    RG_CHECK_ARG(F && K9 && Rt && which && (N == 0 || (pts64 && X)), "null buffers");
    if ((rc = ensure(c->d_in_a, sizeof(double) * 4 * std::max<size_t>(N, 1)))) return rc;
    if ((rc = ensure(c->d_in_b, sizeof(double) * 9 * p))) return rc;
    if ((rc = ensure(c->d_in_c, std::max<size_t>(N, 1)))) return rc;
    if ((rc = ensure(c->d_out_b, sizeof(double) * (12 * p + 3 * std::max<size_t>(N, 1))))) return rc;
    if ((rc = ensure(c->d_out_a, sizeof(int) * p))) return rc;
    if (N) RG_CUDA(cudaMemcpyAsync(c->d_in_a.ptr, pts64, sizeof(double) * 4 * N, cudaMemcpyHostToDevice, st));
    RG_CUDA(cudaMemcpyAsync(c->d_in_b.ptr, F, sizeof(double) * 9 * p, cudaMemcpyHostToDevice, st));
    if (mask && N) RG_CUDA(cudaMemcpyAsync(c->d_in_c.ptr, mask, N, cudaMemcpyHostToDevice, st));
    double* dRt = (double*)c->d_out_b.ptr;
    double* dX = dRt + 12 * p;
    if ((rc = rg_two_view_init_dev(ctx, stream, P, (const double*)c->d_in_a.ptr, pair_off, (const double*)c->d_in_b.ptr, K9,
                                   (mask && N) ? (const unsigned char*)c->d_in_c.ptr : nullptr, dRt, (int32_t*)c->d_out_a.ptr,
                                   dX)))
        return rc;
    RG_CUDA(cudaMemcpyAsync(Rt, dRt, sizeof(double) * 12 * p, cudaMemcpyDeviceToHost, st));
    RG_CUDA(cudaMemcpyAsync(which, c->d_out_a.ptr, sizeof(int) * p, cudaMemcpyDeviceToHost, st));
    if (N) RG_CUDA(cudaMemcpyAsync(X, dX, sizeof(double) * 3 * N, cudaMemcpyDeviceToHost, st));
    RG_CUDA(cudaStreamSynchronize(st));
    return RG_OK;
}

int rg_camera_resectioning_host(void* ctx, void* stream, int V, const double* C, double* K, double* R, double* t) {
    RG_CHECK_ARG(ctx != nullptr, "ctx is null");
    RG_CHECK_ARG(V >= 0, "negative number of cameras");
    Ctx* c = (Ctx*)ctx;
    cudaStream_t st = (cudaStream_t)stream;
    RG_CUDA(cudaSetDevice(c->device));
    if (V == 0) return RG_OK;
    RG_CHECK_ARG(C && K && R && t, "null buffers");
    const size_t v = (size_t)V;
    int rc;
    if ((rc = ensure(c->d_in_b, sizeof(double) * 12 * v))) return rc;
    if ((rc = ensure(c->d_out_b, sizeof(double) * 21 * v))) return rc;
    double* dK = (double*)c->d_out_b.ptr;
    double* dR = dK + 9 * v;
    double* dt = dR + 9 * v;
    RG_CUDA(cudaMemcpyAsync(c->d_in_b.ptr, C, sizeof(double) * 12 * v, cudaMemcpyHostToDevice, st));
    camera_resection_kernel<<<ceil_div(V, 64), 64, 0, st>>>((const double*)c->d_in_b.ptr, V, dK, dR, dt);
    c->last_stats[7] = 1;
    RG_CUDA(cudaGetLastError());
    RG_CUDA(cudaMemcpyAsync(K, dK, sizeof(double) * 9 * v, cudaMemcpyDeviceToHost, st));
    RG_CUDA(cudaMemcpyAsync(R, dR, sizeof(double) * 9 * v, cudaMemcpyDeviceToHost, st));
    RG_CUDA(cudaMemcpyAsync(t, dt, sizeof(double) * 3 * v, cudaMemcpyDeviceToHost, st));
    RG_CUDA(cudaStreamSynchronize(st));
    return RG_OK;
}

// smallest double whose correctly rounded square root is >= tol: sqrt(s) < tol  <=>  s < sqrt_threshold(tol)
static double sqrt_threshold(double tol) {
    double t = tol * tol;
    while (t > 0.0 && std::sqrt(t) >= tol) t = std::nextafter(t, 0.0);
    while (std::sqrt(t) < tol) t = std::nextafter(t, INFINITY);
    return t;
}

int rg_match_first_within_dev(void* ctx, void* stream, int dim, int M, const double* obs_dev, int N, const double* y_dev,
                              double tol, int32_t* idx_dev) {
    RG_CHECK_ARG(ctx != nullptr, "ctx is null");
    RG_CHECK_ARG(dim == 2 || dim == 3, "dim must be 2 or 3");
    RG_CHECK_ARG(M >= 0 && N >= 0, "negative size");
    RG_CHECK_ARG(tol == tol, "tol is NaN");
    Ctx* c = (Ctx*)ctx;
    cudaStream_t st = (cudaStream_t)stream;
    RG_CUDA(cudaSetDevice(c->device));
    c->last_stats[7] = 0;
    if (N == 0) return RG_OK;
    RG_CHECK_ARG(idx_dev && y_dev && (M == 0 || obs_dev), "null buffers");
    RG_CUDA(cudaMemsetAsync(idx_dev, 0xFF, sizeof(int32_t) * (size_t)N, st));           // -1 everywhere
    if (M == 0 || !(tol > 0.0)) return RG_OK;                                            // norm < tol <= 0 never holds
    const double t2 = std::isinf(tol) ? INFINITY : sqrt_threshold(tol);
    // enough (query block x observation segment) blocks to fill the GPU a few times over; segments are whole tiles
    const int qblocks = ceil_div(N, 128);
    int nseg = std::max(1, std::min(ceil_div(M, 256), ceil_div((long long)c->sm_count * 16, qblocks)));
    int seg_len = ceil_div(ceil_div(M, nseg), 256) * 256;
    nseg = ceil_div(M, seg_len);
    RG_CHECK_ARG(nseg <= 65535, "too many observation segments");
    if (dim == 3)
        match_first_kernel<3><<<dim3(qblocks, nseg), 128, 0, st>>>(obs_dev, M, seg_len, y_dev, N, t2, (unsigned*)idx_dev);
    else
        match_first_kernel<2><<<dim3(qblocks, nseg), 128, 0, st>>>(obs_dev, M, seg_len, y_dev, N, t2, (unsigned*)idx_dev);
    c->last_stats[7] += 1;
    RG_CUDA(cudaGetLastError());
    return RG_OK;
}

int rg_match_first_within_host(void* ctx, void* stream, int dim, int M, const double* obs, int N, const double* y,
                               double tol, int32_t* idx) {
    RG_CHECK_ARG(ctx != nullptr, "ctx is null");
    RG_CHECK_ARG(dim == 2 || dim == 3, "dim must be 2 or 3");
    RG_CHECK_ARG(M >= 0 && N >= 0, "negative size");
    Ctx* c = (Ctx*)ctx;
    cudaStream_t st = (cudaStream_t)stream;
    RG_CUDA(cudaSetDevice(c->device));
    if (N == 0) return RG_OK;
    RG_CHECK_ARG(idx && y && (M == 0 || obs), "null buffers");
    int rc;
    const size_t m = (size_t)M, n = (size_t)N, d = (size_t)dim;
    if ((rc = ensure(c->d_in_a, sizeof(double) * d * (m + n) + 16))) return rc;
    if ((rc = ensure(c->d_out_b, sizeof(int) * n))) return rc;
    double* dobs = (double*)c->d_in_a.ptr;
    double* dy = dobs + d * m;
    if (M) RG_CUDA(cudaMemcpyAsync(dobs, obs, sizeof(double) * d * m, cudaMemcpyHostToDevice, st));
    RG_CUDA(cudaMemcpyAsync(dy, y, sizeof(double) * d * n, cudaMemcpyHostToDevice, st));
    if ((rc = rg_match_first_within_dev(ctx, stream, dim, M, dobs, N, dy, tol, (int32_t*)c->d_out_b.ptr))) return rc;
    RG_CUDA(cudaMemcpyAsync(idx, c->d_out_b.ptr, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
    RG_CUDA(cudaStreamSynchronize(st));
    return RG_OK;
}

}  // extern "C"
