// Host-side work planning shared by the F and PnP entry points.
#pragma once
#include "score_core.cuh"
#include <algorithm>
#include <cstdlib>

namespace rg {

struct FPlan {
    int P = 0;
    long long Ntot = 0, Htot = 0, N32tot = 0;
    int n_items = 0;
    int maxN = 0, maxH = 0;
    long long total_words = 0;           // guard-band bitmap size: sum_p H_p * ceil(ngroups_p / 32)
};

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// Fill the PairInfo table (integer part) on the host and upload it.  Work items of the packed scorer:
// item = (pair, 512-hypothesis block, contiguous range of 32-point groups).  The ranges are sized so that the total
// item count is a multiple of the persistent grid when the batch allows it (static round-robin has no tail then).
// resident blocks per SM of a persistent scorer instantiation (queried once; also sets the dynamic smem attribute)
template <class Pol>
inline int score_blocks_per_sm() {
    static int cached = 0;
    if (cached == 0) {
        constexpr size_t smem = score_smem_bytes<Pol>();
        cudaFuncSetAttribute(score_packed<Pol>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, score_packed<Pol>, kScoreThreads, smem) != cudaSuccess || n < 1)
            n = 1;
        cached = n;
    }
    return cached;
}

inline int f_plan(Ctx* c, cudaStream_t st, int P, const int* pair_off, const int* hyp_off, FPlan& plan, int blocks_per_sm,
                  const int* n_vote = nullptr /* PnP: per-view number of voting correspondences (<= view size) */) {
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
        o.n_all = pair_off[p + 1] - pair_off[p];
        o.n = n_vote ? n_vote[p] : o.n_all;
        RG_CHECK_ARG(o.n >= 0 && o.n <= o.n_all, "voting count must be in [0, view size]");
        o.hyp_off = hyp_off[p];
        o.H = hyp_off[p + 1] - hyp_off[p];
        o.n_pad = ((o.n + kSub - 1) / kSub) * kSub;
        o.pt_off32 = (int)off32;
        off32 += o.n_pad;
        plan.maxN = std::max(plan.maxN, o.n);
        plan.maxH = std::max(plan.maxH, o.H);
        const long long nhb = ceil_div(o.H, kHypPerBlock), ng = o.n_pad / kSub;
        unit_total += nhb * ng;
        o.words_per_hyp = (int)((ng + 31) / 32);
        o.word_off = plan.total_words;
        plan.total_words += (long long)o.H * o.words_per_hyp;
    }
    RG_CHECK_ARG(off32 < (1ll << 31) && (long long)hyp_off[P] * 9 < (1ll << 40), "batch too large for 32-bit offsets");
    plan.Ntot = pair_off[P];
    plan.Htot = hyp_off[P];
    plan.N32tot = off32;

    const long long grid = (long long)c->sm_count * blocks_per_sm;
    // aim for ~kItemsPerBlock items per resident block (dynamic scheduling: the tail is about half an item), never below
    // 64 groups (512 points) per item
    static const long long kItemsPerBlock = [] {
        const char* e = getenv("RG_ITEMS_PER_BLOCK");
        const long long v = e ? atoll(e) : 0;
        return v > 0 ? v : 24ll;              // measured on B200, config-5 batch: 8 -> 4.17 ms, 24 -> 4.07 ms, 48 -> 4.07 ms
    }();
    long long gps_target = std::max<long long>(64, unit_total / std::max<long long>(1, grid * kItemsPerBlock));
    long long hb_total = 0;
    for (int p = 0; p < P; ++p) hb_total += ceil_div(pi[p].H, kHypPerBlock) * (pi[p].n_pad > 0 ? 1 : 0);
    // uniform batches: nudge the split so that (#items) % grid == 0
    bool uniform = true;
    for (int p = 1; p < P; ++p) uniform = uniform && pi[p].n_pad == pi[0].n_pad && pi[p].H == pi[0].H;
    if (uniform && hb_total > 0 && pi[0].n_pad > 0) {
        const long long ng = pi[0].n_pad / kSub;
        long long ns0 = std::max<long long>(1, ceil_div(ng, gps_target));
        long long best_ns = ns0;
        // the split that is actually realised after rounding the range to whole bitmap words
        auto realised = [&](long long ns) {
            long long gps = ceil_div(ng, ns);
            if (ns > 1) gps = ((gps + 31) / 32) * 32;
            return (long long)ceil_div(ng, gps);
        };
        for (long long ns = ns0; ns <= std::min<long long>(ng, ns0 * 2 + 4); ++ns) {
            if ((hb_total * realised(ns)) % grid == 0) { best_ns = ns; break; }
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
        if (ns > 1) gps = ((gps + 31) / 32) * 32;      // item boundaries fall on bitmap-word boundaries (1024 points)
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


}  // namespace rg
