// Cross-GPU argmax over NVLink peer memory for the hypothesis-split mode.
//
// Round 1 reduced the 8-byte key with ncclAllReduce(max) and then broadcast the winner's F with a second collective: two
// NCCL launches (~25-40 us on 8 GPUs) next to 0.13 ms of scoring.  Here every rank owns a small buffer that all peers map
// (cudaIpc handles, exchanged once by the host), and ONE kernel per call does the whole exchange:
//   1. store {key, payload (F or R|t)} of my P pairs into MY slot of EVERY peer's buffer (plain stores to peer pointers),
//   2. fence.sys, then store the call's sequence number into my flag word at every peer,
//   3. spin on the local flag words until every rank's sequence number has arrived (bounded: a missing peer raises the
//      status word instead of hanging the GPU),
//   4. take the maximum key per pair (larger count, then LOWER global hypothesis index: the first maximum, fun.py:320-323,
//      ransac.py:108) and copy its payload.
// Slots are double buffered by the parity of the sequence number: a rank that runs ahead writes call k+1 while a slower peer
// still reads call k, and nobody can be two calls ahead because step 3 of call k+1 needs every peer's flag k+1, which a peer
// only writes after finishing call k (stream order).  One rank per GPU — never run several ranks of one exchange on one GPU.
#include "common.cuh"

namespace rg {

constexpr int kP2PMaxWorld = 16;
constexpr int kP2PMaxPairs = 64;
constexpr int kP2PMaxPayload = 12;

struct __align__(16) P2PSlot {
    unsigned long long key;
    unsigned long long pad;
    double payload[kP2PMaxPayload];
};
struct P2PBuf {
    P2PSlot slots[2][kP2PMaxWorld][kP2PMaxPairs];
    unsigned flag[kP2PMaxWorld * 32];               // one 128-byte line per source rank
    unsigned seq;                                   // calls made by THIS rank (only its own kernel touches it): the sequence
                                                    // number lives on the device so that the exchange can sit in a CUDA graph
};
struct P2PPeers { P2PBuf* p[kP2PMaxWorld]; };

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256) p2p_argmax_kernel(P2PBuf* local, P2PPeers peers, int rank, int world, int P,
                                                          const unsigned long long* __restrict__ key,
                                                          const double* __restrict__ payload, int npay,
                                                          int* __restrict__ best_idx, int* __restrict__ best_count,
                                                          double* __restrict__ payload_out, int* __restrict__ status) {
    __shared__ unsigned s_seq;
    if (threadIdx.x == 0) { s_seq = local->seq + 1u; local->seq = s_seq; }
    __syncthreads();
    const unsigned seq = s_seq;
    const int par = (int)(seq & 1u);
    for (int t = threadIdx.x; t < world * P; t += blockDim.x) {
        const int r = t / P, p = t - r * P;
        P2PSlot* dst = &peers.p[r]->slots[par][rank][p];
        dst->key = key[p];
        for (int k = 0; k < npay; ++k) dst->payload[k] = payload[(size_t)p * npay + k];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(&peers.p[threadIdx.x]->flag[rank * 32], seq);
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(&local->flag[threadIdx.x * 32]) - seq) < 0) {
            if (clock64() - t0 > 6000000000ll) { atomicExch(status, 1 + (int)threadIdx.x); break; }    // ~3 s: a peer is missing
        }
        __threadfence_system();
    }
    __syncthreads();
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
        unsigned long long best = 0ull;
        int br = -1;
        for (int r = 0; r < world; ++r) {
            const unsigned long long k = *(volatile const unsigned long long*)&local->slots[par][r][p].key;
            if (k > best) { best = k; br = r; }
        }
        best_count[p] = (int)(best >> 32);
        best_idx[p] = best ? (int)(0xFFFFFFFFu - (unsigned)(best & 0xFFFFFFFFull)) : -1;
        for (int k = 0; k < npay; ++k)
            payload_out[(size_t)p * npay + k] =
                br >= 0 ? *(volatile const double*)&local->slots[par][br][p].payload[k] : __longlong_as_double(0x7FF8000000000000ll);
    }
}

void p2p_release(Ctx* c) {
    for (int r = 0; r < c->p2p_world && r < kP2PMaxWorld; ++r)
        if (c->p2p_peer[r] && r != c->p2p_rank) cudaIpcCloseMemHandle(c->p2p_peer[r]);
    if (c->p2p_local) cudaFree(c->p2p_local);
    for (int r = 0; r < kP2PMaxWorld; ++r) c->p2p_peer[r] = nullptr;
    c->p2p_local = nullptr;
    c->p2p_world = 0;
}

}  // namespace rg

using namespace rg;

extern "C" {

// step 1 (every rank): allocate the local exchange buffer and export its 64-byte IPC handle
int rg_p2p_create(void* ctx, int rank, int world, void* handle_out64) {
    RG_CHECK_ARG(ctx != nullptr && handle_out64 != nullptr, "null argument");
    RG_CHECK_ARG(world >= 1 && world <= kP2PMaxWorld && rank >= 0 && rank < world, "rank / world out of range (world <= 16)");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    Ctx* c = (Ctx*)ctx;
    RG_CUDA(cudaSetDevice(c->device));
    p2p_release(c);
    RG_CUDA(cudaMalloc(&c->p2p_local, sizeof(P2PBuf)));
    RG_CUDA(cudaMemset(c->p2p_local, 0, sizeof(P2PBuf)));
    c->p2p_bytes = sizeof(P2PBuf);
    c->p2p_rank = rank;
    c->p2p_world = world;
    c->p2p_seq = 0;
    c->p2p_peer[rank] = c->p2p_local;
    cudaIpcMemHandle_t h;
    RG_CUDA(cudaIpcGetMemHandle(&h, c->p2p_local));
    memcpy(handle_out64, &h, 64);
    RG_CUDA(cudaDeviceSynchronize());
    return RG_OK;
}

// step 2 (every rank, after the handles were all-gathered by the host): map the peers' buffers
int rg_p2p_connect(void* ctx, const void* handles_all) {
    RG_CHECK_ARG(ctx != nullptr && handles_all != nullptr, "null argument");
    Ctx* c = (Ctx*)ctx;
    RG_CHECK_ARG(c->p2p_local != nullptr, "rg_p2p_create has not been called");
    RG_CUDA(cudaSetDevice(c->device));
    for (int r = 0; r < c->p2p_world; ++r) {
        if (r == c->p2p_rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles_all + 64 * (size_t)r, 64);
        RG_CUDA(cudaIpcOpenMemHandle(&c->p2p_peer[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    return RG_OK;
}

int rg_p2p_destroy(void* ctx) {
    RG_CHECK_ARG(ctx != nullptr, "ctx is null");
    Ctx* c = (Ctx*)ctx;
    RG_CUDA(cudaSetDevice(c->device));
    RG_CUDA(cudaDeviceSynchronize());
    p2p_release(c);
    return RG_OK;
}

// collective: every rank of the exchange calls it the same number of times, in the same order.  key_dev (P) as written by
// rg_f_ransac_dev2 / rg_pnp_ransac_batched_dev2; payload_dev (P x npay doubles, npay <= 12) travels with the winning key.
// status_dev (device int, zero it once): set to 1 + r if rank r did not arrive within ~3 s.
int rg_p2p_argmax_exchange(void* ctx, void* stream, int P, const unsigned long long* key_dev, const double* payload_dev, int npay,
                           int* best_idx_dev, int* best_count_dev, double* payload_out_dev, int* status_dev) {
    RG_CHECK_ARG(ctx != nullptr, "ctx is null");
    Ctx* c = (Ctx*)ctx;
    RG_CHECK_ARG(c->p2p_local != nullptr, "rg_p2p_create / rg_p2p_connect have not been called");
    RG_CHECK_ARG(P >= 0 && P <= kP2PMaxPairs, "at most 64 pairs per exchange");
    RG_CHECK_ARG(npay >= 0 && npay <= kP2PMaxPayload, "payload of at most 12 doubles");
    if (P == 0) return RG_OK;
    RG_CHECK_ARG(key_dev && best_idx_dev && best_count_dev && status_dev && (npay == 0 || (payload_dev && payload_out_dev)),
                 "null buffers");
    for (int r = 0; r < c->p2p_world; ++r) RG_CHECK_ARG(c->p2p_peer[r] != nullptr, "rg_p2p_connect has not mapped every peer");
    RG_CUDA(cudaSetDevice(c->device));
    P2PPeers peers;
    for (int r = 0; r < kP2PMaxWorld; ++r) peers.p[r] = (P2PBuf*)c->p2p_peer[r];
    p2p_argmax_kernel<<<1, 256, 0, (cudaStream_t)stream>>>((P2PBuf*)c->p2p_local, peers, c->p2p_rank, c->p2p_world, P, key_dev,
                                                           payload_dev, npay, best_idx_dev, best_count_dev, payload_out_dev,
                                                           status_dev);
    RG_CUDA(cudaGetLastError());
    return RG_OK;
}

}  // extern "C"
