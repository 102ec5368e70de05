// Host orchestration + C ABI of the F-matrix RANSAC path.  See include/rg_b200.h for the contract of each entry point.
#include "f_kernels.cuh"
#include "jacobi.cuh"
#include "plan.cuh"
#include <algorithm>
#include <vector>

namespace rg {

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

static int f_workspace(Ctx* c, const FPlan& plan) {
    int rc;
    const size_t H = (size_t)std::max<long long>(plan.Htot, 1);
    if ((rc = ensure(c->F64, sizeof(double) * 9 * H))) return rc;
    if ((rc = ensure(c->hyp32, sizeof(Hyp32) * H))) return rc;
    if ((rc = ensure(c->flags, H))) return rc;
    if ((rc = ensure(c->counts, sizeof(int) * (H + 1)))) return rc;      // + the scorer's work counter
    if ((rc = ensure(c->stats, sizeof(unsigned long long) * 8))) return rc;
    if ((rc = ensure(c->best, sizeof(int2) * (size_t)std::max(plan.P, 1)))) return rc;
    if ((rc = ensure(c->tie_stats, sizeof(double2) * H))) return rc;
    if ((rc = ensure_pinned(c->h_stats, sizeof(unsigned long long) * 8))) return rc;
    if ((rc = ensure(c->bitmap, sizeof(unsigned) * (size_t)std::max<long long>(plan.total_words, 1)))) return rc;
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
static int f_score_launch(Ctx* c, cudaStream_t st, const FPlan& plan, const double* pts64, int score_path) {
    PairInfo* pi = (PairInfo*)c->pair_info.ptr;
    int* counts = (int*)c->counts.ptr;
    unsigned long long* stats = (unsigned long long*)c->stats.ptr;
    RG_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * ((size_t)std::max<long long>(plan.Htot, 1) + 1), st));   // counts + work counter
    if (!c->accumulate_stats) RG_CUDA(cudaMemsetAsync(stats, 0, sizeof(unsigned long long) * 8, st));
    if (plan.Htot == 0 || plan.Ntot == 0) return RG_OK;
    if (score_path == SCORE_FP32_GUARDED) {
        if (plan.n_items > 0) {
            constexpr size_t smem = score_smem_bytes<EpiPolicy<MODE>>();
            const int grid = std::min(plan.n_items, c->sm_count * score_blocks_per_sm<EpiPolicy<MODE>>());
            score_packed<EpiPolicy<MODE>><<<grid, kScoreThreads, smem, st>>>(
                (const float4*)c->pts32.ptr, (const Hyp32*)c->hyp32.ptr, pi, plan.P, plan.n_items, counts,
                (unsigned*)c->bitmap.ptr, counts + std::max<long long>(plan.Htot, 1));
            prof_mark(c, st, 3);
            const int fgrid = (int)std::max<long long>(1, std::min<long long>((long long)c->sm_count * 8,
                                                                              (plan.total_words + 255) / 256));
            typename EpiFix<MODE>::Params fp{(const float4*)c->pts32.ptr, (const double4*)pts64, (const Hyp32*)c->hyp32.ptr,
                                             (const double*)c->F64.ptr, pi, plan.P};
            fixup_scan<EpiFix<MODE>><<<fgrid, 256, 0, st>>>(fp, plan.total_words, (const unsigned*)c->bitmap.ptr, counts, stats);
            c->last_stats[7] += 2;
        }
    } else {
        const int zs = std::max(1, std::min(64, ceil_div(plan.maxN, 2048)));
        f_score_fp64<<<dim3(ceil_div(plan.maxH, 128), plan.P, zs), 128, 0, st>>>(
            (const double4*)pts64, (const double*)c->F64.ptr, pi, MODE, counts);
        prof_mark(c, st, 3);
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
    argmax_counts<<<plan.P, 256, 0, st>>>(counts, pi, best, (const unsigned char*)c->flags.ptr,
                                          (unsigned long long*)c->stats.ptr);
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
    int rc = f_plan(c, st, P, pair_off, hyp_off, plan, score_blocks_per_sm<EpiPolicy<MODE_EPI_MAX>>());
    if (rc) return rc;
    for (int p = 0; p < P; ++p) {
        const int n = pair_off[p + 1] - pair_off[p], H = hyp_off[p + 1] - hyp_off[p];
        RG_CHECK_ARG(H == 0 || n >= 8, "a pair with hypotheses needs at least 8 correspondences");
    }
    if ((rc = f_workspace(c, plan))) return rc;
    if (!c->accumulate_stats) c->last_stats[7] = 0;
    if (P == 0) return RG_OK;
    prof_mark(c, st, 0);
    if ((rc = f_prepare(c, st, plan, pts64, thr))) return rc;
    prof_mark(c, st, 1);
    rc = (mode == MODE_SAMPSON) ? f_solve_launch<MODE_SAMPSON>(c, st, plan, pts64, idx, solver)
                                : f_solve_launch<MODE_EPI_MAX>(c, st, plan, pts64, idx, solver);
    if (rc) return rc;
    prof_mark(c, st, 2);
    rc = (mode == MODE_SAMPSON) ? f_score_launch<MODE_SAMPSON>(c, st, plan, pts64, score_path)
                                : f_score_launch<MODE_EPI_MAX>(c, st, plan, pts64, score_path);
    if (rc) return rc;
    prof_mark(c, st, 4);
    if ((rc = f_select_launch(c, st, plan, pts64, mode, tie_mode, mask, best_F, best_idx, best_count))) return rc;
    prof_mark(c, st, 5);
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

// out[0] guard-band groups flagged, [1] band evaluations re-done in FP64, [2] decisions changed by the recheck,
// [7] kernel launches of the last call.  Synchronises the stream.
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

// Host-buffer entry point.  The pairs are processed in up to Ctx::kMaxSlices contiguous sub-batches: all uploads are
// queued on a second stream (one event per sub-batch), the kernels of sub-batch k wait only for their own inputs, so the
// upload of sub-batch k+1 overlaps the scoring of sub-batch k (needs pinned host buffers to actually overlap; pageable
// ones are still correct).  Results are identical to a single batch: pairs are independent.
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
    RG_CHECK_ARG(pair_off[0] == 0 && hyp_off[0] == 0, "offset arrays must start at 0");
    for (int p = 0; p < P; ++p)
        RG_CHECK_ARG(pair_off[p + 1] >= pair_off[p] && hyp_off[p + 1] >= hyp_off[p], "offset arrays must be non-decreasing");
    const size_t Ntot = (size_t)pair_off[P], Htot = (size_t)hyp_off[P];
    RG_CHECK_ARG((Ntot == 0 || pts64) && (Htot == 0 || idx), "input pointers are null");
    int rc;
    if ((rc = ensure(c->d_in_a, sizeof(double) * 4 * std::max<size_t>(Ntot, 1)))) return rc;
    if ((rc = ensure(c->d_in_b, sizeof(int) * 8 * std::max<size_t>(Htot, 1)))) return rc;
    if ((rc = ensure(c->d_out_a, sizeof(int) * 2 * (size_t)P))) return rc;
    if ((rc = ensure(c->d_out_b, sizeof(double) * 9 * (size_t)P))) return rc;
    if (mask && (rc = ensure(c->d_out_c, std::max<size_t>(Ntot, 1)))) return rc;

    // sub-batches.  Every sub-batch pays the small prepare / solve / select kernels again (~0.15 ms), so the automatic choice
    // is two — the first one the smallest whose scoring time covers the upload of everything after it (model: 1.5e12
    // evaluations/s against a 50 GB/s host link), at most half of the pairs — and only when the upload time it hides
    // exceeds that fixed cost (measured on the config-5 batch: 1 -> 5.11, 2 -> 4.80, 3 -> 4.90, 4 -> 4.99 ms; the Dino
    // sequence, 35 small pairs, stays monolithic).
    int first = 1;
    double hidden_s = 0.0;
    {
        double score_s = 0.0;
        for (first = 1; first < P; ++first) {
            const double n = pair_off[first] - pair_off[first - 1], h = hyp_off[first] - hyp_off[first - 1];
            score_s += n * h / 1.5e12;
            const double rest_bytes = 32.0 * ((double)(pair_off[P] - pair_off[first]) + (double)(hyp_off[P] - hyp_off[first]));
            if (score_s >= rest_bytes / 5.0e10) break;
        }
        first = std::max(1, std::min(first, std::max(1, P / 2)));
        double sc = 0.0;
        for (int q = 0; q < first && q < P; ++q)
            sc += (double)(pair_off[q + 1] - pair_off[q]) * (double)(hyp_off[q + 1] - hyp_off[q]) / 1.5e12;
        const double rest_bytes = P > 0 ? 32.0 * ((double)(pair_off[P] - pair_off[std::min(first, P)]) +
                                                  (double)(hyp_off[P] - hyp_off[std::min(first, P)])) : 0.0;
        hidden_s = std::min(sc, rest_bytes / 5.0e10);
    }
    int S = c->opt_host_slices > 0 ? c->opt_host_slices : (hidden_s > 1.5e-4 ? 2 : 1);
    S = std::max(1, std::min(std::min(S, P), (int)Ctx::kMaxSlices));
    if (c->opt_profile) S = 1;                       // phase events describe one monolithic call
    if (S > 1) {
        if (!c->copy_stream) RG_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        if (!c->copy_gate) RG_CUDA(cudaEventCreateWithFlags(&c->copy_gate, cudaEventDisableTiming));
        for (int k = 0; k < S; ++k)
            if (!c->slice_ready[k]) RG_CUDA(cudaEventCreateWithFlags(&c->slice_ready[k], cudaEventDisableTiming));
    }
    // the first sub-batch is small (its upload is the only one that nothing hides), the others share the rest evenly
    int bounds[Ctx::kMaxSlices + 1];
    bounds[0] = 0;
    if (S == 1) {
        bounds[1] = P;
    } else {
        for (int k = 1; k <= S; ++k) bounds[k] = first + (int)((long long)(P - first) * (k - 1) / (S - 1));
    }
    double* d_pts = (double*)c->d_in_a.ptr;
    int* d_ix = (int*)c->d_in_b.ptr;
    cudaStream_t cs = S > 1 ? c->copy_stream : st;
    if (S > 1) {                                      // uploads may not overtake earlier work queued on the caller's stream
        RG_CUDA(cudaEventRecord(c->copy_gate, st));
        RG_CUDA(cudaStreamWaitEvent(cs, c->copy_gate, 0));
    }
    for (int k = 0; k < S; ++k) {
        const int p0 = bounds[k], p1 = bounds[k + 1];
        const size_t n0 = (size_t)pair_off[p0], n1 = (size_t)pair_off[p1], h0 = (size_t)hyp_off[p0], h1 = (size_t)hyp_off[p1];
        if (n1 > n0) RG_CUDA(cudaMemcpyAsync(d_pts + 4 * n0, pts64 + 4 * n0, sizeof(double) * 4 * (n1 - n0), cudaMemcpyHostToDevice, cs));
        if (h1 > h0) RG_CUDA(cudaMemcpyAsync(d_ix + 8 * h0, idx + 8 * h0, sizeof(int) * 8 * (h1 - h0), cudaMemcpyHostToDevice, cs));
        if (S > 1) RG_CUDA(cudaEventRecord(c->slice_ready[k], cs));
    }
    int* d_idx = (int*)c->d_out_a.ptr;
    int* d_cnt = d_idx + P;
    std::vector<int> po, ho;
    rc = RG_OK;
    for (int k = 0; k < S && rc == RG_OK; ++k) {
        const int p0 = bounds[k], p1 = bounds[k + 1], Pk = p1 - p0;
        if (Pk == 0) continue;
        const size_t n0 = (size_t)pair_off[p0], h0 = (size_t)hyp_off[p0];
        const int* pok = pair_off;
        const int* hok = hyp_off;
        if (k > 0 || S > 1) {                          // offsets relative to the sub-batch
            po.resize(Pk + 1); ho.resize(Pk + 1);
            for (int q = 0; q <= Pk; ++q) { po[q] = pair_off[p0 + q] - (int)n0; ho[q] = hyp_off[p0 + q] - (int)h0; }
            pok = po.data(); hok = ho.data();
        }
        if (S > 1) RG_CUDA(cudaStreamWaitEvent(st, c->slice_ready[k], 0));
        c->accumulate_stats = k > 0;
        rc = f_ransac_dev(c, st, Pk, d_pts + 4 * n0, pok, d_ix + 8 * h0, hok, thr, mode, tie_mode, solver, score_path,
                          d_idx + p0, d_cnt + p0, (double*)c->d_out_b.ptr + 9 * (size_t)p0,
                          mask ? (unsigned char*)c->d_out_c.ptr + n0 : nullptr);
        c->accumulate_stats = false;
        if (rc) break;
        // per-hypothesis results live in per-call workspaces: fetch them before the next sub-batch overwrites them
        const size_t Hk = (size_t)hyp_off[p1] - h0;
        if (counts && Hk) RG_CUDA(cudaMemcpyAsync(counts + h0, c->counts.ptr, sizeof(int) * Hk, cudaMemcpyDeviceToHost, st));
        if (F_all && Hk) RG_CUDA(cudaMemcpyAsync(F_all + 9 * h0, c->F64.ptr, sizeof(double) * 9 * Hk, cudaMemcpyDeviceToHost, st));
        if (flags && Hk) RG_CUDA(cudaMemcpyAsync(flags + h0, c->flags.ptr, Hk, cudaMemcpyDeviceToHost, st));
    }
    if (rc) { cudaStreamSynchronize(st); if (S > 1) cudaStreamSynchronize(cs); return rc; }
    RG_CUDA(cudaMemcpyAsync(best_idx, d_idx, sizeof(int) * (size_t)P, cudaMemcpyDeviceToHost, st));
    RG_CUDA(cudaMemcpyAsync(best_count, d_cnt, sizeof(int) * (size_t)P, cudaMemcpyDeviceToHost, st));
    RG_CUDA(cudaMemcpyAsync(best_F, c->d_out_b.ptr, sizeof(double) * 9 * (size_t)P, cudaMemcpyDeviceToHost, st));
    if (mask && Ntot) RG_CUDA(cudaMemcpyAsync(mask, c->d_out_c.ptr, Ntot, cudaMemcpyDeviceToHost, st));
    RG_CUDA(cudaStreamSynchronize(st));
    // sample indices are validated where they are read (the solver clamps them and sets flag bit 2): no host pass over idx
    if (c->h_stats.ptr && ((const unsigned long long*)c->h_stats.ptr)[4] != 0) {
        set_error("invalid argument: %llu hypotheses have a sample index outside [0, N) of their pair",
                  ((const unsigned long long*)c->h_stats.ptr)[4]);
        return RG_ERR_ARG;
    }
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
    int rc = f_plan(c, st, 1, pair_off, hyp_off, plan, score_blocks_per_sm<EpiPolicy<MODE_EPI_MAX>>());
    if (rc) return rc;
    if ((rc = f_workspace(c, plan))) return rc;
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
    rc = (mode == MODE_SAMPSON) ? f_score_launch<MODE_SAMPSON>(c, st, plan, (const double*)c->d_in_a.ptr, score_path)
                                : f_score_launch<MODE_EPI_MAX>(c, st, plan, (const double*)c->d_in_a.ptr, score_path);
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
    int rc = f_plan(c, st, 1, pair_off, hyp_off, plan, score_blocks_per_sm<EpiPolicy<MODE_EPI_MAX>>());
    if (rc) return rc;
    if ((rc = f_workspace(c, plan))) return rc;
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
