// Generic persistent FP32 scorer skeleton shared by the F path (epipolar criterion) and the PnP path (reprojection
// criterion).  A policy supplies the packed point layout, the hypothesis record and the two-point evaluation.
#pragma once
#include "common.cuh"

namespace rg {

#ifndef RG_SCORE_BLOCKS
#define RG_SCORE_BLOCKS 3
#endif
#ifndef RG_EPI_CHUNK
#define RG_EPI_CHUNK 512
#endif
#ifndef RG_SCORE_THREADS
#define RG_SCORE_THREADS 256
#endif
#ifndef RG_HPT
#define RG_HPT 2
#endif
#ifndef RG_PAIR_UNROLL
#define RG_PAIR_UNROLL 4          // point pairs of a flag group evaluated per unrolled loop body (kSub / 2 = all)
#endif
constexpr int kScoreThreads = RG_SCORE_THREADS;
constexpr int kScoreBlocksPerSM = RG_SCORE_BLOCKS;
constexpr int kHypPerThread = RG_HPT;
constexpr int kHypPerBlock  = kScoreThreads * kHypPerThread;   // 512 hypotheses per work item
constexpr int kStages       = 2;
constexpr int kPairUnroll   = RG_PAIR_UNROLL;

// packed FMA with scalar operands broadcast inside the instruction (SASS: FFMA2 Rd, Rs.F32, Rb.F32x2, ...).  Building the
// {s, s} pair inside the asm block keeps the coefficient in ONE register: a duplicated pair would cost a second
// register-file read per use, and the scorer is bound by register-file read ports ("scorer").
__device__ __forceinline__ float2 ffma2_sbc(float s, float2 b, float2 c) {          // {s,s} * b + c
    float2 d;
    asm("{\n.reg .b64 ts, tb, tc, td;\nmov.b64 ts, {%2, %2};\nmov.b64 tb, {%3, %4};\nmov.b64 tc, {%5, %6};\n"
        "fma.rn.f32x2 td, ts, tb, tc;\nmov.b64 {%0, %1}, td;\n}"
        : "=f"(d.x), "=f"(d.y) : "f"(s), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ float2 ffma2_sbs(float s, float2 b, float c) {           // {s,s} * b + {c,c}
    float2 d;
    asm("{\n.reg .b64 ts, tb, tc, td;\nmov.b64 ts, {%2, %2};\nmov.b64 tb, {%3, %4};\nmov.b64 tc, {%5, %5};\n"
        "fma.rn.f32x2 td, ts, tb, tc;\nmov.b64 {%0, %1}, td;\n}"
        : "=f"(d.x), "=f"(d.y) : "f"(s), "f"(b.x), "f"(b.y), "f"(c));
    return d;
}

struct ScoreItem {
    int pair, h_base, H_end;        // hypotheses [h_base, min(h_base + kHypPerBlock, H_end)) (global indices)
    int g0, g1;                     // kSub-point groups [g0, g1) of the pair (g0 is a multiple of 32)
    int W;                          // bitmap words per hypothesis of this pair
    long long wbase;                // bitmap word of (first hypothesis of the item, group g0)
};

__device__ __forceinline__ ScoreItem decode_item(const PairInfo* __restrict__ pi, int P, int item) {
    int lo = 0, hi = P;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (pi[mid].item_off <= item) lo = mid; else hi = mid; }
    const PairInfo& info = pi[lo];
    const int local = item - info.item_off;
    const int hb = local / info.nsplit;
    const int sp = local - hb * info.nsplit;
    ScoreItem it;
    it.pair = lo;
    it.h_base = info.hyp_off + hb * kHypPerBlock;
    it.H_end = info.hyp_off + info.H;
    const int ngroups = info.n_pad / kSub;
    it.g0 = min(sp * info.groups_per_split, ngroups);
    it.g1 = min(it.g0 + info.groups_per_split, ngroups);
    it.W = info.words_per_hyp;
    it.wbase = info.word_off + (long long)(hb * kHypPerBlock) * info.words_per_hyp + (it.g0 >> 5);
    return it;
}

// Guard-band bookkeeping: one bit per (hypothesis, kSub-point group).  A set bit means "the FP32 result of at least
// one evaluation of this group lies inside the rounding band of this hypothesis" and makes the fix-up kernel re-evaluate
// that group in FP64.  Every 32-bit word (32 groups = 256 points of one hypothesis) is written by exactly one thread of
// exactly one work item, so the scorer needs no atomics and the bitmap needs no clearing.
// Policy requirements:
//   typedef Rec;                       hypothesis record in global/shared memory (sizeof % 16 == 0)
//   typedef Regs;                      hypothesis in registers
//   static constexpr int kVec4PerPair; float4 per packed point pair
//   static constexpr int kChunkPts;    points per shared-memory stage
//   static void load(const Rec* sh, int slot, bool valid, Regs&, float& G);
//   template <int K> static void evalN(const Regs (&H)[K], const float4* pair, unsigned (&cnt)[K], float (&minabs)[K]);
template <class Pol>
struct __align__(128) ScoreStage {
    float4 pts[Pol::kChunkPts / 2 * Pol::kVec4PerPair];
    typename Pol::Rec hyp[kHypPerBlock];
};
template <class Pol>
constexpr size_t score_smem_bytes() { return kStages * sizeof(ScoreStage<Pol>) + 64; }

// persistent block: first item blockIdx.x, then items claimed from a global counter (dynamic: with static round-robin a
// batch whose item count is not a multiple of the grid leaves most SMs idle during the last round); every item = 512
// hypotheses x a contiguous range of kSub-point groups of one pair.  Host guarantees every item has >= 1 group and >= 1
// hypothesis; *work_counter is 0 at launch.
// Points and (at the first chunk of an item) hypothesis records arrive by 1-D bulk TMA on one mbarrier per stage;
// the next chunk / next item is always in flight while the current one is being scored.
template <class Pol>
__global__ void __launch_bounds__(kScoreThreads, kScoreBlocksPerSM)
score_packed(const float4* __restrict__ pts32, const typename Pol::Rec* __restrict__ hyp32,
             const PairInfo* __restrict__ pi, int P, int n_items, int* __restrict__ counts,
             unsigned* __restrict__ bitmap, int* __restrict__ work_counter) {
    __shared__ int s_next;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ScoreStage<Pol>* st = reinterpret_cast<ScoreStage<Pol>*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + kStages * sizeof(ScoreStage<Pol>));
    constexpr int kChunkPts = Pol::kChunkPts;
    constexpr uint32_t kPtBytes = Pol::kVec4PerPair * 8;        // bytes per point in the packed layout
    constexpr int kV = Pol::kVec4PerPair;

    const int tid = threadIdx.x;
    int item = blockIdx.x;
    if (item >= n_items) return;
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_mbar_init();
    }
    __syncthreads();

    uint32_t phases = 0u;
    int stage = 0;

    ScoreItem cur = decode_item(pi, P, item);
    // packed points of the pair: kV float4 per point pair
    const float4* cur_src = pts32 + (size_t)(pi[cur.pair].pt_off32 / 2) * kV;
    if (tid == 0) {
        const uint32_t nb = (uint32_t)min((cur.g1 - cur.g0) * kSub, kChunkPts) * kPtBytes;
        const uint32_t hb = (uint32_t)min(kHypPerBlock, cur.H_end - cur.h_base) * (uint32_t)sizeof(typename Pol::Rec);
        mbar_expect_tx(&full[0], nb + hb);
        tma_load_1d(st[0].pts, cur_src + (size_t)cur.g0 * (kSub / 2) * kV, nb, &full[0]);
        tma_load_1d(st[0].hyp, hyp32 + cur.h_base, hb, &full[0]);
    }

    while (true) {
        // claim the following item now: its first chunk and hypotheses are prefetched during the last chunk of this one
        if (tid == 0) s_next = (int)gridDim.x + atomicAdd(work_counter, 1);
        __syncthreads();                  // (the previous value was read by everybody before the last chunk's barrier)
        const int next_item = s_next;
        const bool has_next = next_item < n_items;
        ScoreItem nxt = cur;
        const float4* nxt_src = cur_src;
        if (has_next) {
            nxt = decode_item(pi, P, next_item);
            nxt_src = pts32 + (size_t)(pi[nxt.pair].pt_off32 / 2) * kV;
        }
        typename Pol::Regs Hy[kHypPerThread];
        float G[kHypPerThread];
        unsigned cnt[kHypPerThread];
        int hid[kHypPerThread];
        unsigned* wptr[kHypPerThread];
#pragma unroll
        for (int k = 0; k < kHypPerThread; ++k) {
            G[k] = 0.f; cnt[k] = 0u;
            hid[k] = cur.h_base + k * kScoreThreads + tid;
            wptr[k] = bitmap + cur.wbase + (long long)(k * kScoreThreads + tid) * cur.W;
        }

        const int total_pts = (cur.g1 - cur.g0) * kSub;
        for (int done = 0; done < total_pts; done += kChunkPts) {
            const int npts = min(total_pts - done, kChunkPts);
            // prefetch the following chunk (of this item, or the first chunk + hypotheses of the next item)
            if (tid == 0) {
                const int s2 = stage ^ 1;
                if (done + kChunkPts < total_pts) {
                    const uint32_t nb = (uint32_t)min(total_pts - done - kChunkPts, kChunkPts) * kPtBytes;
                    mbar_expect_tx(&full[s2], nb);
                    tma_load_1d(st[s2].pts, cur_src + (size_t)(cur.g0 * kSub + done + kChunkPts) / 2 * kV, nb, &full[s2]);
                } else if (has_next) {
                    const uint32_t nb = (uint32_t)min((nxt.g1 - nxt.g0) * kSub, kChunkPts) * kPtBytes;
                    const uint32_t hb =
                        (uint32_t)min(kHypPerBlock, nxt.H_end - nxt.h_base) * (uint32_t)sizeof(typename Pol::Rec);
                    mbar_expect_tx(&full[s2], nb + hb);
                    tma_load_1d(st[s2].pts, nxt_src + (size_t)nxt.g0 * (kSub / 2) * kV, nb, &full[s2]);
                    tma_load_1d(st[s2].hyp, hyp32 + nxt.h_base, hb, &full[s2]);
                }
            }
            mbar_wait(&full[stage], (phases >> stage) & 1u);
            phases ^= 1u << stage;

            if (done == 0) {
#pragma unroll
                for (int k = 0; k < kHypPerThread; ++k)
                    Pol::load(st[stage].hyp, k * kScoreThreads + tid, hid[k] < cur.H_end, Hy[k], G[k]);
            }
            const float4* sp = st[stage].pts;
            // chunk = whole bitmap words (32 flag groups of kSub points each), except the last word of an item
            const int ngr = npts / kSub;
            const int wfirst = done / (kSub * 32);
            for (int wq = 0; wq * 32 < ngr; ++wq) {
                const int ng = min(32, ngr - wq * 32);
                unsigned flag[kHypPerThread];
#pragma unroll
                for (int k = 0; k < kHypPerThread; ++k) flag[k] = 0u;
                const float4* wp = sp + wq * 32 * (kSub / 2) * kV;
                for (int g = 0; g < ng; ++g) {
                    float ma[kHypPerThread];
#pragma unroll
                    for (int k = 0; k < kHypPerThread; ++k) ma[k] = INFINITY;
                    const float4* gp = wp + g * (kSub / 2) * kV;
#pragma unroll kPairUnroll
                    for (int j = 0; j < kSub / 2; ++j) Pol::template evalN<kHypPerThread>(Hy, gp + j * kV, cnt, ma);
#pragma unroll
                    for (int k = 0; k < kHypPerThread; ++k) flag[k] |= (ma[k] <= G[k] ? 1u : 0u) << g;
                }
#pragma unroll
                for (int k = 0; k < kHypPerThread; ++k)
                    if (hid[k] < cur.H_end) wptr[k][wfirst + wq] = flag[k];
            }
            __syncthreads();          // everyone is done with this stage before it is refilled
            stage ^= 1;
        }
#pragma unroll
        for (int k = 0; k < kHypPerThread; ++k)
            if (hid[k] < cur.H_end && cnt[k]) atomicAdd(&counts[hid[k]], (int)cnt[k]);
        if (!has_next) break;
        item = next_item;
        cur = nxt;
        cur_src = nxt_src;
    }
}

// FP64 fix-up.  Two levels of warp compaction keep all 32 lanes busy although only ~0.4 % of the bitmap bits are set
// and only ~1 in 8 evaluations of a flagged group is inside the band:
//   level 1  scan the bitmap (coalesced, 4 words per lane in flight), queue the set bits; one lane per flagged group
//            re-evaluates its kSub correspondences in FP32 (bit-identical op sequence) -> band mask + FP32 decisions;
//   level 2  queue the band evaluations; one lane per evaluation applies the reference formula in FP64 and corrects
//            counts[h] by (exact decision - FP32 decision).
// stats: [0] flagged groups, [1] band evaluations, [2] changed decisions.
//   Fix::decode(params, word_index, h, flag_base, aux)            hypothesis / first flag index / pair of a bitmap word
//   Fix::scan(params, h, flag, aux, band, sign)                   FP32 pass over one group (bit k = correspondence k)
//   Fix::exact(params, h, i, aux)                                 FP64 decision for correspondence i of the pair/view
template <class Fix>
__global__ void __launch_bounds__(256) fixup_scan(typename Fix::Params prm, long long total_words,
                                                   const unsigned* __restrict__ bitmap, int* __restrict__ counts,
                                                   unsigned long long* __restrict__ stats) {
    __shared__ int4 queue[8][64];
    __shared__ int4 queue2[8][64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int4* q = queue[warp];
    int4* q2 = queue2[warp];
    int qn = 0, q2n = 0;
    unsigned long long n_groups = 0, n_band = 0, n_flip = 0;
    const unsigned lt = (1u << lane) - 1u;
    const unsigned full = 0xffffffffu;

    auto drain2 = [&](int n) {                      // lanes < n: one band evaluation each, FP64
        if (lane < n) {
            const int4 e = q2[lane];
            const int d = Fix::exact(prm, e.x, e.y, e.z) - e.w;
            if (d) { atomicAdd(&counts[e.x], d); n_flip += 1; }
            n_band += 1;
        }
        __syncwarp();
    };
    auto push2 = [&](bool has, int4 rec) {          // warp-collective append to the level-2 queue
        const unsigned m = __ballot_sync(full, has);
        if (has) q2[q2n + __popc(m & lt)] = rec;
        q2n += __popc(m);
        __syncwarp();
        if (q2n >= 32) {
            drain2(32);
            const int rest = q2n - 32;
            int4 mv = make_int4(0, 0, 0, 0);
            if (lane < rest) mv = q2[32 + lane];
            __syncwarp();
            if (lane < rest) q2[lane] = mv;
            __syncwarp();
            q2n = rest;
        }
    };
    auto drain = [&](int n) {                       // lanes < n: one flagged group each, FP32 re-evaluation
        unsigned band = 0u, sign = 0u;
        int4 r = make_int4(0, 0, 0, 0);
        if (lane < n) {
            r = q[lane];
            Fix::scan(prm, r.x, r.y, r.z, band, sign);
            n_groups += 1;
        }
        __syncwarp();
        while (__any_sync(full, band != 0u)) {
            const bool has = band != 0u;
            int k = 0;
            if (has) { k = __ffs(band) - 1; band &= band - 1; }
            push2(has, make_int4(r.x, r.y * kSub + k, r.z, (int)((sign >> k) & 1u)));
        }
    };

    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long wbase = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); wbase < total_words; wbase += 4 * stride) {
        unsigned words[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {               // four independent loads in flight per lane
            const long long wi = wbase + u * stride + lane;
            words[u] = (wi < total_words) ? bitmap[wi] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            unsigned word = words[u];
            if (!__any_sync(full, word != 0u)) continue;
            int h = 0, fbase = 0, aux = 0;
            if (word) Fix::decode(prm, wbase + u * stride + lane, h, fbase, aux);
            while (__any_sync(full, word != 0u)) {
                const bool has = word != 0u;
                const unsigned m = __ballot_sync(full, has);
                if (has) {
                    const int b = __ffs(word) - 1;
                    word &= word - 1;
                    q[qn + __popc(m & lt)] = make_int4(h, fbase + b, aux, 0);
                }
                qn += __popc(m);
                __syncwarp();
                if (qn >= 32) {
                    drain(32);
                    const int rest = qn - 32;
                    int4 mv = make_int4(0, 0, 0, 0);
                    if (lane < rest) mv = q[32 + lane];
                    __syncwarp();
                    if (lane < rest) q[lane] = mv;
                    __syncwarp();
                    qn = rest;
                }
            }
        }
    }
    drain(qn);
    drain2(q2n);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_groups += __shfl_xor_sync(full, n_groups, o);
        n_band += __shfl_xor_sync(full, n_band, o);
        n_flip += __shfl_xor_sync(full, n_flip, o);
    }
    if (lane == 0) {
        if (n_groups) atomicAdd(&stats[0], n_groups);
        if (n_band) atomicAdd(&stats[1], n_band);
        if (n_flip) atomicAdd(&stats[2], n_flip);
    }
}

// best[p] = {index inside the pair of the first hypothesis with the largest count (-1 if that count is 0), count}
// flags / stats (optional): hypotheses whose flag byte has bit 2 set (sample index out of range) are counted in stats[4]
__global__ void __launch_bounds__(256) argmax_counts(const int* __restrict__ counts, const PairInfo* __restrict__ pi,
                                                      int2* __restrict__ best, const unsigned char* __restrict__ flags,
                                                      unsigned long long* __restrict__ stats) {
    __shared__ unsigned long long sk[8];
    const int p = blockIdx.x;
    const PairInfo info = pi[p];
    unsigned long long key = 0ull;
    unsigned bad = 0u;
    for (int h = threadIdx.x; h < info.H; h += blockDim.x) {
        const unsigned long long k =
            ((unsigned long long)(unsigned)counts[info.hyp_off + h] << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)h);
        key = k > key ? k : key;
        if (flags != nullptr) bad += (flags[info.hyp_off + h] >> 2) & 1u;
    }
    if (stats != nullptr && bad) atomicAdd(&stats[4], (unsigned long long)bad);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other > key ? other : key;
    }
    if ((threadIdx.x & 31) == 0) sk[threadIdx.x >> 5] = key;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) key = sk[w] > key ? sk[w] : key;
        const int cnt = (int)(key >> 32);
        const int idx = cnt > 0 ? (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull)) : -1;
        best[p] = make_int2(idx, cnt);
    }
}

}  // namespace rg
