// Cross-GPU argmax for a pair / view whose hypotheses are split over several GPUs: every rank
// scores its hypothesis block, packs (inlier count, global hypothesis index) into one monotone 64-bit key on the device
// and ONE ncclAllReduce(max) of 8 bytes per pair picks the winner — larger count first, then the LOWER index, i.e. the
// first maximum, as the single-GPU selection (fun.py:320-323, ransac.py:108).  NCCL is resolved with dlopen at the first
// call, so librg_b200.so has no link-time dependency on it and single-GPU users never load it.
#include "common.cuh"
#include <dlfcn.h>

namespace rg {

typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*nccl_errstr_fn)(int);
constexpr int kNcclUint64 = 5;      // ncclDataType_t::ncclUint64 (nccl.h)
constexpr int kNcclMax = 2;         // ncclRedOp_t::ncclMax

static nccl_allreduce_fn g_allreduce = nullptr;
static nccl_errstr_fn g_errstr = nullptr;

static int load_nccl() {
    if (g_allreduce) return RG_OK;
    const char* names[] = {getenv("RG_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* nm : names) {
        if (!nm || !*nm) continue;
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        set_error("NCCL not found (dlopen libnccl.so.2: %s); set RG_NCCL_LIB", dlerror());
        return RG_ERR_ARG;
    }
    g_allreduce = (nccl_allreduce_fn)dlsym(h, "ncclAllReduce");
    g_errstr = (nccl_errstr_fn)dlsym(h, "ncclGetErrorString");
    if (!g_allreduce) {
        set_error("ncclAllReduce not found in the NCCL library");
        return RG_ERR_ARG;
    }
    return RG_OK;
}

// key = (count << 32) | (0xFFFFFFFF - global index); 0 when the rank has no hypothesis with an inlier
__global__ void argmax_pack_keys(const int* __restrict__ best_idx, const int* __restrict__ best_count, int P, int index_offset,
                                 unsigned long long* __restrict__ key) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const int i = best_idx[p], c = best_count[p];
    key[p] = (i >= 0 && c > 0)
                 ? (((unsigned long long)(unsigned)c << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)(i + index_offset)))
                 : 0ull;
}

__global__ void argmax_unpack_keys(const unsigned long long* __restrict__ key, int P, int* __restrict__ best_idx,
                                   int* __restrict__ best_count) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const unsigned long long k = key[p];
    best_count[p] = (int)(k >> 32);
    best_idx[p] = k ? (int)(0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull)) : -1;
}

}  // namespace rg

using namespace rg;

extern "C" {

int rg_argmax_pack_dev(void* stream, int P, const int32_t* best_idx_dev, const int32_t* best_count_dev, int index_offset,
                       unsigned long long* key_dev) {
    RG_CHECK_ARG(P >= 0, "negative size");
    if (P == 0) return RG_OK;
    RG_CHECK_ARG(best_idx_dev && best_count_dev && key_dev, "null buffers");
    argmax_pack_keys<<<(P + 127) / 128, 128, 0, (cudaStream_t)stream>>>(best_idx_dev, best_count_dev, P, index_offset, key_dev);
    RG_CUDA(cudaGetLastError());
    return RG_OK;
}

int rg_argmax_unpack_dev(void* stream, int P, const unsigned long long* key_dev, int32_t* best_idx_dev, int32_t* best_count_dev) {
    RG_CHECK_ARG(P >= 0, "negative size");
    if (P == 0) return RG_OK;
    RG_CHECK_ARG(best_idx_dev && best_count_dev && key_dev, "null buffers");
    argmax_unpack_keys<<<(P + 127) / 128, 128, 0, (cudaStream_t)stream>>>(key_dev, P, best_idx_dev, best_count_dev);
    RG_CUDA(cudaGetLastError());
    return RG_OK;
}

int rg_argmax_allreduce(void* nccl_comm, void* stream, unsigned long long* key_dev, int count) {
    RG_CHECK_ARG(nccl_comm != nullptr, "nccl_comm is null");
    RG_CHECK_ARG(count >= 0, "negative size");
    if (count == 0) return RG_OK;
    RG_CHECK_ARG(key_dev != nullptr, "key_dev is null");
    int rc = load_nccl();
    if (rc) return rc;
    const int r = g_allreduce(key_dev, key_dev, (size_t)count, kNcclUint64, kNcclMax, nccl_comm, (cudaStream_t)stream);
    if (r != 0) {
        set_error("ncclAllReduce failed: %s", g_errstr ? g_errstr(r) : "unknown NCCL error");
        return RG_ERR_CUDA;
    }
    return RG_OK;
}

}  // extern "C"
