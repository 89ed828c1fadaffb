// gather.cu — the two exchanges of the multi-GPU paths behind the C ABI, for host programs that shard
// work over one process per GPU without Python: the all-gather of per-pair registration results (batched pair ICP,
// config C5) and the all-gather + merge of per-shard loop-closure candidates (config C4).  Both are KB-sized, once
// per batch / per detect(): latency, not bandwidth.
//
// NCCL is not linked: the caller owns the communicator (ncclComm_t, passed as void*) and therefore already has NCCL
// loaded; the two entry points needed are looked up in that library at first use (dlopen with RTLD_NOLOAD first, so
// that the communicator and ncclAllGather come from the same library instance).
#include <dlfcn.h>

#include <algorithm>

#include "common.cuh"

namespace sb {

typedef int (*nccl_all_gather_fn)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef const char* (*nccl_error_fn)(int);
static nccl_all_gather_fn g_all_gather = nullptr;
static nccl_error_fn g_error = nullptr;
static constexpr int NCCL_UINT8 = 1;   // ncclUint8 (nccl.h: ncclInt8 = 0, ncclUint8 = 1, ...): stable since NCCL 2.0

static int nccl_bind(Ctx* ctx) {
    if (g_all_gather) return SB_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(ctx, SB_ERR_INVALID_ARG, "gather: libnccl.so.2 is not loaded and cannot be found (%s)", dlerror());
    g_all_gather = reinterpret_cast<nccl_all_gather_fn>(dlsym(h, "ncclAllGather"));
    g_error = reinterpret_cast<nccl_error_fn>(dlsym(h, "ncclGetErrorString"));
    if (!g_all_gather) return fail(ctx, SB_ERR_INVALID_ARG, "gather: ncclAllGather not found in libnccl");
    return SB_OK;
}

// every rank contributes `bytes` bytes (the same on all ranks); out: world * bytes, rank-major.  Host in, host out.
static int all_gather_bytes(Ctx* ctx, void* comm, int world, const void* local, size_t bytes, void* out) {
    SB_TRY(nccl_bind(ctx));
    char *d_send, *d_recv;
    SB_TRY(arena_get(ctx, bytes ? bytes : 1, &d_send));
    SB_TRY(arena_get(ctx, bytes * (size_t)world + 1, &d_recv));
    if (bytes) SB_CUDA(ctx, cudaMemcpyAsync(d_send, local, bytes, cudaMemcpyHostToDevice, ctx->stream));
    const int r = g_all_gather(d_send, d_recv, bytes, NCCL_UINT8, comm, ctx->stream);
    if (r != 0) return fail(ctx, SB_ERR_CUDA, "gather: ncclAllGather failed: %s", g_error ? g_error(r) : "?");
    if (bytes) SB_CUDA(ctx, cudaMemcpyAsync(out, d_recv, bytes * (size_t)world, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}

}  // namespace sb

using namespace sb;

extern "C" {

int sb_gather_results(sb_ctx* ctx, void* nccl_comm, int32_t rank, int32_t world, const sb_icp_result* local,
                      int32_t n_total, sb_icp_result* all) {
    if (!ctx || !nccl_comm || world < 1 || rank < 0 || rank >= world || n_total < 0 || (n_total > 0 && !all))
        return SB_ERR_INVALID_ARG;
    Ctx* c = &ctx->c;
    cudaSetDevice(c->device);
    SB_TRY(arena_reset(c));
    const int per = (n_total + world - 1) / world;                 // units of the fullest rank
    const int mine = n_total > rank ? (n_total - rank + world - 1) / world : 0;   // units u = rank, rank + world, ...
    if (mine > 0 && !local) return fail(c, SB_ERR_INVALID_ARG, "gather: rank %d owns %d results but `local` is null", rank, mine);
    if (per == 0) return SB_OK;
    std::vector<sb_icp_result> send((size_t)per), recv((size_t)per * (size_t)world);
    memset(send.data(), 0, sizeof(sb_icp_result) * send.size());
    for (int i = 0; i < mine; ++i) send[(size_t)i] = local[i];
    SB_TRY(all_gather_bytes(c, nccl_comm, world, send.data(), sizeof(sb_icp_result) * (size_t)per, recv.data()));
    for (int u = 0; u < n_total; ++u) all[u] = recv[(size_t)(u % world) * (size_t)per + (size_t)(u / world)];
    return SB_OK;
}

int sb_gather_candidates(sb_ctx* ctx, void* nccl_comm, int32_t world, const double* dist, const int32_t* entry,
                         int32_t n_local, int32_t capacity, double* out_dist, int32_t* out_entry, int32_t* out_count) {
    if (!ctx || !nccl_comm || world < 1 || n_local < 0 || capacity < 1 || !out_dist || !out_entry || !out_count ||
        (n_local > 0 && (!dist || !entry)))
        return SB_ERR_INVALID_ARG;
    Ctx* c = &ctx->c;
    cudaSetDevice(c->device);
    SB_TRY(arena_reset(c));
    struct Rec { double d; long long e; };   // e = -1: padding
    std::vector<Rec> send((size_t)capacity), recv((size_t)capacity * (size_t)world);
    for (int i = 0; i < capacity; ++i) {
        send[(size_t)i].d = i < n_local ? dist[i] : 0.0;
        send[(size_t)i].e = i < n_local ? (long long)entry[i] : -1;
    }
    SB_TRY(all_gather_bytes(c, nccl_comm, world, send.data(), sizeof(Rec) * (size_t)capacity, recv.data()));
    std::vector<std::pair<double, int>> m;
    for (const Rec& r : recv)
        if (r.e >= 0) m.push_back({r.d, (int)r.e});
    std::sort(m.begin(), m.end());   // loop_closure.hpp:92: ascending (distance, entry), identical on every rank
    const int n = (int)std::min(m.size(), (size_t)capacity);
    for (int i = 0; i < n; ++i) { out_dist[i] = m[(size_t)i].first; out_entry[i] = m[(size_t)i].second; }
    *out_count = n;
    return SB_OK;
}

}  // extern "C"
