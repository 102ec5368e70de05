// Host orchestration + C ABI of the F-matrix RANSAC path.  See include/rg_b200.h for the contract of each entry point.
#include "f_kernels.cuh"
#include "jacobi.cuh"
#include <algorithm>
#include <vector>

namespace rg {

struct FPlan {
    int P = 0;
    long long Ntot = 0, Htot = 0, N32tot = 0;
    int n_items = 0;
    int maxN = 0, maxH = 0;
    long long total_groups_hyps = 0;     // sum_p H_p * ngroups_p  (upper bound on recheck records)
};

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// Fill the PairInfo table (integer part) on the host and upload it.  Work items of the packed scorer:
// item = (pair, 512-hypothesis block, contiguous range of 32-point groups).  The ranges are sized so that the total
// item count is a multiple of the persistent grid when the batch allows it (static round-robin has no tail then).
static int f_plan(Ctx* c, cudaStream_t st, int P, const int* pair_off, const int* hyp_off, FPlan& plan) {
    RG_CHECK_ARG(P >= 0 && P <= 65535, "number of pairs must be in [0, 65535]");
    RG_CHECK_ARG(pair_off != nullptr && hyp_off != nullptr, "offset arrays are null");
    RG_CHECK_ARG(pair_off[0] == 0 && hyp_off[0] == 0, "offset arrays must start at 0");
    plan = FPlan();
    plan.P = P;
    if (P == 0) return RG_OK;
    for (int p = 0; p < P; ++p) {
        RG_CHECK_ARG(pair_off[p + 1] >= pair_off[p] && hyp_off[p + 1] >= hyp_off[p], "offset arrays must be non-decreasing");
    }
    int rc = ensure_pinned(c->h_stage, sizeof(PairInfo) * (size_t)P);
    if (rc) return rc;
    RG_CUDA(cudaEventSynchronize(c->staging_free));
    PairInfo* pi = (PairInfo*)c->h_stage.ptr;

    long long unit_total = 0;      // work in units of (hypothesis block x point group)
    long long off32 = 0;
    for (int p = 0; p < P; ++p) {
        PairInfo& o = pi[p];
        memset(&o, 0, sizeof(o));
        o.pt_off = pair_off[p];
        o.n = pair_off[p + 1] - pair_off[p];
        o.hyp_off = hyp_off[p];
        o.H = hyp_off[p + 1] - hyp_off[p];
        o.n_pad = ((o.n + kSub - 1) / kSub) * kSub;
        o.pt_off32 = (int)off32;
        off32 += o.n_pad;
        plan.maxN = std::max(plan.maxN, o.n);
        plan.maxH = std::max(plan.maxH, o.H);
        const long long nhb = ceil_div(o.H, kHypPerBlock), ng = o.n_pad / kSub;
        unit_total += nhb * ng;
        plan.total_groups_hyps += (long long)o.H * ng;
    }
    RG_CHECK_ARG(off32 < (1ll << 31) && (long long)hyp_off[P] * 9 < (1ll << 40), "batch too large for 32-bit offsets");
    plan.Ntot = pair_off[P];
    plan.Htot = hyp_off[P];
    plan.N32tot = off32;

    const long long grid = (long long)c->sm_count * 2;
    // aim for ~8 items per resident block, never below 8 groups (256 points) per item
    long long gps_target = std::max<long long>(8, unit_total / std::max<long long>(1, grid * 8));
    long long hb_total = 0;
    for (int p = 0; p < P; ++p) hb_total += ceil_div(pi[p].H, kHypPerBlock) * (pi[p].n_pad > 0 ? 1 : 0);
    // uniform batches: nudge the split so that (#items) % grid == 0
    bool uniform = true;
    for (int p = 1; p < P; ++p) uniform = uniform && pi[p].n_pad == pi[0].n_pad && pi[p].H == pi[0].H;
    if (uniform && hb_total > 0 && pi[0].n_pad > 0) {
        const long long ng = pi[0].n_pad / kSub;
        long long ns0 = std::max<long long>(1, ceil_div(ng, gps_target));
        long long best_ns = ns0;
        for (long long ns = ns0; ns <= std::min<long long>(ng, ns0 * 2 + 4); ++ns) {
            if ((hb_total * ns) % grid == 0) { best_ns = ns; break; }
        }
        gps_target = std::max<long long>(1, ceil_div(ng, best_ns));
    }
    long long item_off = 0;
    for (int p = 0; p < P; ++p) {
        PairInfo& o = pi[p];
        const int ng = o.n_pad / kSub;
        const int nhb = ceil_div(o.H, kHypPerBlock);
        o.item_off = (int)item_off;
        if (ng == 0 || nhb == 0) { o.nsplit = 1; o.groups_per_split = 0; continue; }     // contributes no items
        int gps = (int)std::min<long long>(ng, gps_target);
        int ns = ceil_div(ng, gps);
        gps = ceil_div(ng, ns);
        ns = ceil_div(ng, gps);
        o.nsplit = ns;
        o.groups_per_split = gps;
        item_off += (long long)nhb * ns;
    }
    RG_CHECK_ARG(item_off < (1ll << 31), "too many scorer work items");
    plan.n_items = (int)item_off;
    // (pairs without items share their item_off with the next pair; decode_item takes the LAST pair whose
    //  item_off <= item, which is never an empty one)
    rc = ensure(c->pair_info, sizeof(PairInfo) * (size_t)P);
    if (rc) return rc;
    RG_CUDA(cudaMemcpyAsync(c->pair_info.ptr, pi, sizeof(PairInfo) * (size_t)P, cudaMemcpyHostToDevice, st));
    RG_CUDA(cudaEventRecord(c->staging_free, st));
    return RG_OK;
}

static int f_prepare(Ctx* c, cudaStream_t st, const FPlan& plan, const double* pts64, double thr) {
    const int P = plan.P;
    int rc;
    if ((rc = ensure(c->bbox, sizeof(int) * 8 * (size_t)P))) return rc;
    if ((rc = ensure(c->pts32, sizeof(float4) * (size_t)std::max<long long>(plan.N32tot, 1)))) return rc;
    PairInfo* pi = (PairInfo*)c->pair_info.ptr;
    f_bbox_init<<<ceil_div(P * 8, 256), 256, 0, st>>>((int*)c->bbox.ptr, P);
    const int nbx = std::max(1, std::min(64, ceil_div(plan.maxN, 256 * 4)));
    f_bbox<<<dim3(nbx, P), 256, 0, st>>>((const double4*)pts64, pi, (int*)c->bbox.ptr);
    f_frame<<<ceil_div(P, 128), 128, 0, st>>>(pi, (const int*)c->bbox.ptr, P, thr);
    const int nbn = std::max(1, std::min(128, ceil_div(plan.maxN / 2 + kSub, 256)));
    f_normalise<<<dim3(nbn, P), 256, 0, st>>>((const double4*)pts64, pi, (float4*)c->pts32.ptr);
    c->last_stats[7] += 4;
    RG_CUDA(cudaGetLastError());
    return RG_OK;
}

static int f_workspace(Ctx* c, const FPlan& plan, int* wl_cap) {
    int rc;
    const size_t H = (size_t)std::max<long long>(plan.Htot, 1);
    if ((rc = ensure(c->F64, sizeof(double) * 9 * H))) return rc;
    if ((rc = ensure(c->hyp32, sizeof(Hyp32) * H))) return rc;
    if ((rc = ensure(c->flags, H))) return rc;
    if ((rc = ensure(c->counts, sizeof(int) * H))) return rc;
    if ((rc = ensure(c->stats, sizeof(unsigned long long) * 8))) return rc;
    if ((rc = ensure(c->best, sizeof(int2) * (size_t)std::max(plan.P, 1)))) return rc;
    if ((rc = ensure(c->tie_stats, sizeof(double2) * H))) return rc;
    if ((rc = ensure_pinned(c->h_stats, sizeof(unsigned long long) * 8))) return rc;
    long long cap = std::min<long long>(plan.total_groups_hyps, std::max<long long>(1ll << 22, plan.total_groups_hyps / 8));
    cap = std::max<long long>(cap, 1024);
    cap = std::min<long long>(cap, (1ll << 31) - 1);
    if ((rc = ensure(c->worklist, sizeof(int2) * (size_t)cap))) return rc;
    *wl_cap = (int)cap;
    return RG_OK;
}

template <int MODE>
static int f_solve_launch(Ctx* c, cudaStream_t st, const FPlan& plan, const double* pts64, const int* idx, int solver) {
    PairInfo* pi = (PairInfo*)c->pair_info.ptr;
    if (plan.Htot == 0) return RG_OK;
    if (solver == SOLVER_QR) {
        f8_solve_qr<MODE><<<ceil_div(plan.Htot, 128), 128, 0, st>>>((const double4*)pts64, idx, pi, plan.P, (int)plan.Htot,
                                                                  (double*)c->F64.ptr, (Hyp32*)c->hyp32.ptr,
                                                                  (unsigned char*)c->flags.ptr);
    } else {
        const int groups_per_block = kJacobiThreads / 16;
        f8_solve_jacobi<MODE><<<ceil_div(plan.Htot, groups_per_block), kJacobiThreads, 0, st>>>(
            (const double4*)pts64, idx, pi, plan.P, (int)plan.Htot, (double*)c->F64.ptr, (Hyp32*)c->hyp32.ptr,
            (unsigned char*)c->flags.ptr);
    }
    c->last_stats[7] += 1;
    RG_CUDA(cudaGetLastError());
    return RG_OK;
}

template <int MODE>
static int f_score_launch(Ctx* c, cudaStream_t st, const FPlan& plan, const double* pts64, int score_path, int wl_cap) {
    PairInfo* pi = (PairInfo*)c->pair_info.ptr;
    int* counts = (int*)c->counts.ptr;
    unsigned long long* stats = (unsigned long long*)c->stats.ptr;
    RG_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * (size_t)std::max<long long>(plan.Htot, 1), st));
    RG_CUDA(cudaMemsetAsync(stats, 0, sizeof(unsigned long long) * 8, st));
    if (plan.Htot == 0 || plan.Ntot == 0) return RG_OK;
    if (score_path == SCORE_FP32_GUARDED) {
        if (plan.n_items > 0) {
            static bool attr_set = false;
            if (!attr_set) {
                RG_CUDA(cudaFuncSetAttribute(f_score_packed<MODE_EPI_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)kScoreSmemBytes));
                RG_CUDA(cudaFuncSetAttribute(f_score_packed<MODE_SAMPSON>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)kScoreSmemBytes));
                attr_set = true;
            }
            const int grid = std::min(plan.n_items, c->sm_count * 2);
            f_score_packed<MODE><<<grid, kScoreThreads, kScoreSmemBytes, st>>>(
                (const float4*)c->pts32.ptr, (const Hyp32*)c->hyp32.ptr, pi, plan.P, plan.n_items, counts,
                (int2*)c->worklist.ptr, stats, wl_cap);
            f_fixup<MODE><<<c->sm_count * 4, 256, 0, st>>>((const float4*)c->pts32.ptr, (const double4*)pts64,
                                                           (const Hyp32*)c->hyp32.ptr, (const double*)c->F64.ptr, pi, plan.P,
                                                           counts, (const int2*)c->worklist.ptr, stats, wl_cap);
            // exact FP64 rescoring of everything, executed only if the recheck work-list overflowed (stats[3] != 0)
            f_score_fp64<<<dim3(ceil_div(plan.maxH, 128), plan.P, 1), 128, 0, st>>>(
                (const double4*)pts64, (const double*)c->F64.ptr, pi, MODE, counts, stats + 3);
            c->last_stats[7] += 3;
        }
    } else {
        const int zs = std::max(1, std::min(64, ceil_div(plan.maxN, 2048)));
        f_score_fp64<<<dim3(ceil_div(plan.maxH, 128), plan.P, zs), 128, 0, st>>>(
            (const double4*)pts64, (const double*)c->F64.ptr, pi, MODE, counts, nullptr);
        c->last_stats[7] += 1;
    }
    RG_CUDA(cudaGetLastError());
    return RG_OK;
}

static int f_select_launch(Ctx* c, cudaStream_t st, const FPlan& plan, const double* pts64, int mode, int tie_mode,
                           unsigned char* mask, double* best_F, int* best_idx, int* best_count) {
    if (plan.P == 0) return RG_OK;
    PairInfo* pi = (PairInfo*)c->pair_info.ptr;
    int* counts = (int*)c->counts.ptr;
    int2* best = (int2*)c->best.ptr;
    f_argmax<<<plan.P, 256, 0, st>>>(counts, pi, best);
    c->last_stats[7] += 1;
    if (tie_mode == TIE_REFERENCE && plan.Htot > 0) {
        const int nb = (int)std::min<long long>(plan.Htot, (long long)c->sm_count * 8);
        f_tie_stats<<<nb, 256, 0, st>>>((const double4*)pts64, (const double*)c->F64.ptr, counts, pi, plan.P,
                                        (int)plan.Htot, best, mode, (double2*)c->tie_stats.ptr);
        f_tie_resolve<<<plan.P, 32, 0, st>>>(counts, pi, (const double2*)c->tie_stats.ptr, best);
        c->last_stats[7] += 2;
    }
    const int nbx = std::max(1, std::min(64, ceil_div(plan.maxN, 256 * 4)));
    f_mask<<<dim3(nbx, plan.P), 256, 0, st>>>((const double4*)pts64, (const double*)c->F64.ptr, pi, best, mode, mask, best_F,
                                              best_idx, best_count);
    c->last_stats[7] += 1;
    RG_CUDA(cudaGetLastError());
    return RG_OK;
}

static int f_stats_readback(Ctx* c, cudaStream_t st) {
    RG_CUDA(cudaMemcpyAsync(c->h_stats.ptr, c->stats.ptr, sizeof(unsigned long long) * 8, cudaMemcpyDeviceToHost, st));
    return RG_OK;
}

// full pipeline on device pointers; asynchronous with respect to the host except for the tiny PairInfo staging
static int f_ransac_dev(Ctx* c, cudaStream_t st, int P, const double* pts64, const int* pair_off, const int* idx,
                        const int* hyp_off, double thr, int mode, int tie_mode, int solver, int score_path, int* best_idx,
                        int* best_count, double* best_F, unsigned char* mask) {
    RG_CHECK_ARG(thr > 0.0 && std::isfinite(thr), "thr must be positive and finite");
    RG_CHECK_ARG(mode == MODE_EPI_MAX || mode == MODE_SAMPSON, "unknown scoring mode");
    RG_CHECK_ARG(tie_mode == TIE_FIRST || tie_mode == TIE_REFERENCE, "unknown tie mode");
    RG_CHECK_ARG(solver == SOLVER_QR || solver == SOLVER_JACOBI, "unknown solver");
    RG_CHECK_ARG(score_path == SCORE_FP32_GUARDED || score_path == SCORE_FP64, "unknown scoring path");
    RG_CUDA(cudaSetDevice(c->device));
    FPlan plan;
    int rc = f_plan(c, st, P, pair_off, hyp_off, plan);
    if (rc) return rc;
    for (int p = 0; p < P; ++p) {
        const int n = pair_off[p + 1] - pair_off[p], H = hyp_off[p + 1] - hyp_off[p];
        RG_CHECK_ARG(H == 0 || n >= 8, "a pair with hypotheses needs at least 8 correspondences");
    }
    int wl_cap = 0;
    if ((rc = f_workspace(c, plan, &wl_cap))) return rc;
    c->last_stats[7] = 0;
    if (P == 0) return RG_OK;
    if ((rc = f_prepare(c, st, plan, pts64, thr))) return rc;
    rc = (mode == MODE_SAMPSON) ? f_solve_launch<MODE_SAMPSON>(c, st, plan, pts64, idx, solver)
                                : f_solve_launch<MODE_EPI_MAX>(c, st, plan, pts64, idx, solver);
    if (rc) return rc;
    rc = (mode == MODE_SAMPSON) ? f_score_launch<MODE_SAMPSON>(c, st, plan, pts64, score_path, wl_cap)
                                : f_score_launch<MODE_EPI_MAX>(c, st, plan, pts64, score_path, wl_cap);
    if (rc) return rc;
    if ((rc = f_select_launch(c, st, plan, pts64, mode, tie_mode, mask, best_F, best_idx, best_count))) return rc;
    return f_stats_readback(c, st);
}

}  // namespace rg

using namespace rg;

extern "C" {

int rg_f_ransac_dev(void* ctx, void* stream, int P, const double* pts64_dev, const int* pair_off_host, const int* idx_dev,
                    const int* hyp_off_host, double thr, int mode, int tie_mode, int solver, int score_path,
                    int* best_idx_dev, int* best_count_dev, double* best_F_dev, unsigned char* mask_dev) {
    RG_CHECK_ARG(ctx != nullptr, "ctx is null");
    RG_CHECK_ARG(best_idx_dev && best_count_dev && best_F_dev, "output pointers are null");
    return f_ransac_dev((Ctx*)ctx, (cudaStream_t)stream, P, pts64_dev, pair_off_host, idx_dev, hyp_off_host, thr, mode,
                        tie_mode, solver, score_path, best_idx_dev, best_count_dev, best_F_dev, mask_dev);
}

// device pointers to the per-hypothesis results of the LAST call on this context (valid until the next call)
int rg_f_last_hypotheses_dev(void* ctx, const int** counts_dev, const double** F_all_dev, const unsigned char** flags_dev) {
    RG_CHECK_ARG(ctx != nullptr, "ctx is null");
    Ctx* c = (Ctx*)ctx;
    if (counts_dev) *counts_dev = (const int*)c->counts.ptr;
    if (F_all_dev) *F_all_dev = (const double*)c->F64.ptr;
    if (flags_dev) *flags_dev = (const unsigned char*)c->flags.ptr;
    return RG_OK;
}

// out[0] recheck groups pushed, [1] band evaluations re-done in FP64, [2] decisions changed by the recheck,
// [3] work-list overflow (FP64 rescoring was executed), [7] kernel launches of the last call.  Synchronises the stream.
int rg_get_last_stats(void* ctx, void* stream, long long* out8) {
    RG_CHECK_ARG(ctx != nullptr && out8 != nullptr, "null argument");
    Ctx* c = (Ctx*)ctx;
    RG_CUDA(cudaSetDevice(c->device));
    RG_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    for (int i = 0; i < 8; ++i) out8[i] = 0;
    if (c->h_stats.ptr) {
        const unsigned long long* s = (const unsigned long long*)c->h_stats.ptr;
        for (int i = 0; i < 7; ++i) out8[i] = (long long)s[i];
    }
    out8[7] = c->last_stats[7];
    return RG_OK;
}

int rg_f_ransac_host(void* ctx, void* stream, int P, const double* pts64, const int* pair_off, const int* idx,
                     const int* hyp_off, double thr, int mode, int tie_mode, int solver, int score_path, int* best_idx,
                     int* best_count, double* best_F, unsigned char* mask, int* counts, double* F_all, unsigned char* flags) {
    RG_CHECK_ARG(ctx != nullptr, "ctx is null");
    RG_CHECK_ARG(P >= 0 && pair_off && hyp_off, "bad pair table");
    RG_CHECK_ARG(best_idx && best_count && best_F, "output pointers are null");
    Ctx* c = (Ctx*)ctx;
    cudaStream_t st = (cudaStream_t)stream;
    RG_CUDA(cudaSetDevice(c->device));
    if (P == 0) return RG_OK;
    const size_t Ntot = (size_t)pair_off[P], Htot = (size_t)hyp_off[P];
    RG_CHECK_ARG((Ntot == 0 || pts64) && (Htot == 0 || idx), "input pointers are null");
    int rc;
    if ((rc = ensure(c->d_in_a, sizeof(double) * 4 * std::max<size_t>(Ntot, 1)))) return rc;
    if ((rc = ensure(c->d_in_b, sizeof(int) * 8 * std::max<size_t>(Htot, 1)))) return rc;
    if ((rc = ensure(c->d_out_a, sizeof(int) * 2 * (size_t)P))) return rc;
    if ((rc = ensure(c->d_out_b, sizeof(double) * 9 * (size_t)P))) return rc;
    if (mask && (rc = ensure(c->d_out_c, std::max<size_t>(Ntot, 1)))) return rc;
    if (Ntot) RG_CUDA(cudaMemcpyAsync(c->d_in_a.ptr, pts64, sizeof(double) * 4 * Ntot, cudaMemcpyHostToDevice, st));
    if (Htot) RG_CUDA(cudaMemcpyAsync(c->d_in_b.ptr, idx, sizeof(int) * 8 * Htot, cudaMemcpyHostToDevice, st));
    int* d_idx = (int*)c->d_out_a.ptr;
    int* d_cnt = d_idx + P;
    rc = f_ransac_dev(c, st, P, (const double*)c->d_in_a.ptr, pair_off, (const int*)c->d_in_b.ptr, hyp_off, thr, mode,
                      tie_mode, solver, score_path, d_idx, d_cnt, (double*)c->d_out_b.ptr,
                      mask ? (unsigned char*)c->d_out_c.ptr : nullptr);
    if (rc) return rc;
    RG_CUDA(cudaMemcpyAsync(best_idx, d_idx, sizeof(int) * (size_t)P, cudaMemcpyDeviceToHost, st));
    RG_CUDA(cudaMemcpyAsync(best_count, d_cnt, sizeof(int) * (size_t)P, cudaMemcpyDeviceToHost, st));
    RG_CUDA(cudaMemcpyAsync(best_F, c->d_out_b.ptr, sizeof(double) * 9 * (size_t)P, cudaMemcpyDeviceToHost, st));
    if (mask && Ntot) RG_CUDA(cudaMemcpyAsync(mask, c->d_out_c.ptr, Ntot, cudaMemcpyDeviceToHost, st));
    if (counts && Htot) RG_CUDA(cudaMemcpyAsync(counts, c->counts.ptr, sizeof(int) * Htot, cudaMemcpyDeviceToHost, st));
    if (F_all && Htot) RG_CUDA(cudaMemcpyAsync(F_all, c->F64.ptr, sizeof(double) * 9 * Htot, cudaMemcpyDeviceToHost, st));
    if (flags && Htot) RG_CUDA(cudaMemcpyAsync(flags, c->flags.ptr, Htot, cudaMemcpyDeviceToHost, st));
    RG_CUDA(cudaStreamSynchronize(st));
    return RG_OK;
}

// Stage entry point: score caller-supplied fundamental matrices (H x 9, row-major, pixel frame) on one pair.
int rg_epi_score_count_host(void* ctx, void* stream, int N, const double* pts64, int H, const double* F_all, double thr,
                            int mode, int score_path, int* counts) {
    RG_CHECK_ARG(ctx != nullptr, "ctx is null");
    RG_CHECK_ARG(N >= 0 && H >= 0 && counts, "bad sizes / null output");
    RG_CHECK_ARG(thr > 0.0 && std::isfinite(thr), "thr must be positive and finite");
    RG_CHECK_ARG(mode == MODE_EPI_MAX || mode == MODE_SAMPSON, "unknown scoring mode");
    Ctx* c = (Ctx*)ctx;
    cudaStream_t st = (cudaStream_t)stream;
    RG_CUDA(cudaSetDevice(c->device));
    if (H == 0) return RG_OK;
    const int pair_off[2] = {0, N}, hyp_off[2] = {0, H};
    FPlan plan;
    int rc = f_plan(c, st, 1, pair_off, hyp_off, plan);
    if (rc) return rc;
    int wl_cap = 0;
    if ((rc = f_workspace(c, plan, &wl_cap))) return rc;
    if ((rc = ensure(c->d_in_a, sizeof(double) * 4 * std::max<size_t>((size_t)N, 1)))) return rc;
    if (N) RG_CUDA(cudaMemcpyAsync(c->d_in_a.ptr, pts64, sizeof(double) * 4 * (size_t)N, cudaMemcpyHostToDevice, st));
    RG_CUDA(cudaMemcpyAsync(c->F64.ptr, F_all, sizeof(double) * 9 * (size_t)H, cudaMemcpyHostToDevice, st));
    c->last_stats[7] = 0;
    if ((rc = f_prepare(c, st, plan, (const double*)c->d_in_a.ptr, thr))) return rc;
    PairInfo* pi = (PairInfo*)c->pair_info.ptr;
    if (mode == MODE_SAMPSON)
        f_make_hyp32<MODE_SAMPSON><<<ceil_div(H, 256), 256, 0, st>>>((const double*)c->F64.ptr, pi, 1, H, (Hyp32*)c->hyp32.ptr);
    else
        f_make_hyp32<MODE_EPI_MAX><<<ceil_div(H, 256), 256, 0, st>>>((const double*)c->F64.ptr, pi, 1, H, (Hyp32*)c->hyp32.ptr);
    c->last_stats[7] += 1;
    rc = (mode == MODE_SAMPSON) ? f_score_launch<MODE_SAMPSON>(c, st, plan, (const double*)c->d_in_a.ptr, score_path, wl_cap)
                                : f_score_launch<MODE_EPI_MAX>(c, st, plan, (const double*)c->d_in_a.ptr, score_path, wl_cap);
    if (rc) return rc;
    if ((rc = f_stats_readback(c, st))) return rc;
    RG_CUDA(cudaMemcpyAsync(counts, c->counts.ptr, sizeof(int) * (size_t)H, cudaMemcpyDeviceToHost, st));
    RG_CUDA(cudaStreamSynchronize(st));
    return RG_OK;
}

// Stage entry point: 8-point solve only.  idx is (H, 8) indices into the N points.
int rg_f8pt_solve_host(void* ctx, void* stream, int N, const double* pts64, int H, const int* idx, int solver, double* F_all,
                       unsigned char* flags) {
    RG_CHECK_ARG(ctx != nullptr, "ctx is null");
    RG_CHECK_ARG(N >= 8 && H >= 0 && pts64 && F_all, "need N >= 8 and non-null buffers");
    RG_CHECK_ARG(solver == SOLVER_QR || solver == SOLVER_JACOBI, "unknown solver");
    Ctx* c = (Ctx*)ctx;
    cudaStream_t st = (cudaStream_t)stream;
    RG_CUDA(cudaSetDevice(c->device));
    if (H == 0) return RG_OK;
    RG_CHECK_ARG(idx != nullptr, "idx is null");
    const int pair_off[2] = {0, N}, hyp_off[2] = {0, H};
    FPlan plan;
    int rc = f_plan(c, st, 1, pair_off, hyp_off, plan);
    if (rc) return rc;
    int wl_cap = 0;
    if ((rc = f_workspace(c, plan, &wl_cap))) return rc;
    if ((rc = ensure(c->d_in_a, sizeof(double) * 4 * (size_t)N))) return rc;
    if ((rc = ensure(c->d_in_b, sizeof(int) * 8 * (size_t)H))) return rc;
    RG_CUDA(cudaMemcpyAsync(c->d_in_a.ptr, pts64, sizeof(double) * 4 * (size_t)N, cudaMemcpyHostToDevice, st));
    RG_CUDA(cudaMemcpyAsync(c->d_in_b.ptr, idx, sizeof(int) * 8 * (size_t)H, cudaMemcpyHostToDevice, st));
    c->last_stats[7] = 0;
    if ((rc = f_prepare(c, st, plan, (const double*)c->d_in_a.ptr, 1.0))) return rc;
    if ((rc = f_solve_launch<MODE_EPI_MAX>(c, st, plan, (const double*)c->d_in_a.ptr, (const int*)c->d_in_b.ptr, solver)))
        return rc;
    RG_CUDA(cudaMemcpyAsync(F_all, c->F64.ptr, sizeof(double) * 9 * (size_t)H, cudaMemcpyDeviceToHost, st));
    if (flags) RG_CUDA(cudaMemcpyAsync(flags, c->flags.ptr, (size_t)H, cudaMemcpyDeviceToHost, st));
    RG_CUDA(cudaStreamSynchronize(st));
    return RG_OK;
}

// lab3.fmatrix_residuals drop-in arithmetic: x, y are (2, N) row-major; out is (2, N) row-major.
int rg_fmatrix_residuals_host(void* ctx, void* stream, const double* F9, int N, const double* x, const double* y, double* out) {
    RG_CHECK_ARG(ctx != nullptr, "ctx is null");
    RG_CHECK_ARG(F9 && N >= 0, "bad arguments");
    Ctx* c = (Ctx*)ctx;
    cudaStream_t st = (cudaStream_t)stream;
    RG_CUDA(cudaSetDevice(c->device));
    if (N == 0) return RG_OK;
    RG_CHECK_ARG(x && y && out, "null buffers");
    int rc;
    if ((rc = ensure(c->d_in_a, sizeof(double) * (4 * (size_t)N + 16)))) return rc;
    if ((rc = ensure(c->d_out_b, sizeof(double) * 2 * (size_t)N))) return rc;
    double* d = (double*)c->d_in_a.ptr;
    RG_CUDA(cudaMemcpyAsync(d, F9, sizeof(double) * 9, cudaMemcpyHostToDevice, st));
    RG_CUDA(cudaMemcpyAsync(d + 16, x, sizeof(double) * 2 * (size_t)N, cudaMemcpyHostToDevice, st));
    RG_CUDA(cudaMemcpyAsync(d + 16 + 2 * (size_t)N, y, sizeof(double) * 2 * (size_t)N, cudaMemcpyHostToDevice, st));
    f_residuals<<<std::max(1, std::min(c->sm_count * 4, ceil_div(N, 256))), 256, 0, st>>>(d, d + 16, d + 16 + 2 * (size_t)N, N,
                                                                                         (double*)c->d_out_b.ptr);
    RG_CUDA(cudaGetLastError());
    RG_CUDA(cudaMemcpyAsync(out, c->d_out_b.ptr, sizeof(double) * 2 * (size_t)N, cudaMemcpyDeviceToHost, st));
    RG_CUDA(cudaStreamSynchronize(st));
    return RG_OK;
}

}  // extern "C"
