// shard.cu — screen-space row-band sharding across the GPUs of one box (no reference counterpart: the
// reference is a single-GPU sample).
//
// One process per GPU.  A pass image of height H is split into `world` contiguous bands of whole 8-row tiles
// (hr_shard_rows).  Every rank keeps a full replica of the G-buffer, the BVH and the blue-noise tables, computes all
// stages on its band plus a recompute halo (rays and stencils are pure functions of replicated inputs, so the halo is
// exact), and after the last stage of a pass the ranks exchange their bands of
//   * the pass's final output (the denoised frame every rank ends up with), and
//   * the temporal history surfaces the next frame's reprojection gathers from (prev_image / moments, AO colour /
//     history length), because reprojection may read any row under camera motion,
// with one NCCL group of per-band broadcasts (an all-gather with unequal counts: 2160 rows = 270 tiles do not divide
// evenly by 8).  After the exchange every rank holds exactly the images a single GPU would hold: the 1/2/4/8-GPU
// results are bit-identical (tests/test_gpu_multi.py).
//
// NCCL is dlopen'ed on first use so the single-GPU library has no NCCL dependency; inside a PyTorch process the
// already-loaded torch-bundled libnccl.so.2 is picked up.
#include "hr_internal.h"
#include <dlfcn.h>
#include <nccl.h>

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*)                                                                 = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int)                                          = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t)                                                                    = nullptr;
    ncclResult_t (*GroupStart)()                                                                               = nullptr;
    ncclResult_t (*GroupEnd)()                                                                                 = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t)       = nullptr;
    const char* (*GetErrorString)(ncclResult_t)                                                                = nullptr;
    bool ok = false;
};

NcclApi& nccl()
{
    static NcclApi api;
    if (api.handle) return api;
    const char* names[] = { "libnccl.so.2", "libnccl.so" };
    for (const char* n : names)
    {
        api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) return api;
#define LOAD(field, sym) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, sym))
    LOAD(GetUniqueId, "ncclGetUniqueId");
    LOAD(CommInitRank, "ncclCommInitRank");
    LOAD(CommDestroy, "ncclCommDestroy");
    LOAD(GroupStart, "ncclGroupStart");
    LOAD(GroupEnd, "ncclGroupEnd");
    LOAD(Broadcast, "ncclBroadcast");
    LOAD(GetErrorString, "ncclGetErrorString");
#undef LOAD
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.GroupStart && api.GroupEnd && api.Broadcast && api.GetErrorString;
    return api;
}

} // namespace

#define HR_NCCL(ctx, expr)                                                                                      \
    do {                                                                                                        \
        ncclResult_t _r = (expr);                                                                               \
        if (_r != ncclSuccess) {                                                                                \
            hr_set_error((ctx), "%s failed: %s (%s:%d)", #expr, nccl().GetErrorString(_r), __FILE__, __LINE__); \
            return HR_ERR_NCCL;                                                                                 \
        }                                                                                                       \
    } while (0)

void hr_band(const hr_ctx* ctx, int H, int* b0, int* b1)
{
    if (ctx->world <= 1) { *b0 = 0; *b1 = H; return; }
    hr_shard_rows(H, ctx->rank, ctx->world, b0, b1);
}

void hr_wait_exchange(hr_pass* p, cudaStream_t st)
{
    if (p && p->xchg_pending && p->ev_done) cudaStreamWaitEvent(st, p->ev_done, 0);
}

int hr_shard_exchange(hr_pass* p, const ExchangeItem* items, int n, cudaStream_t caller)
{
    hr_ctx* ctx = p->ctx;
    if (ctx->world <= 1 || !ctx->nccl_comm) return HR_OK;
    NcclApi& N = nccl();
    if (!ctx->comm_stream) HR_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
    if (!p->ev_ready)
    {
        HR_CUDA(ctx, cudaEventCreateWithFlags(&p->ev_ready, cudaEventDisableTiming));
        HR_CUDA(ctx, cudaEventCreateWithFlags(&p->ev_done, cudaEventDisableTiming));
    }
    HR_CUDA(ctx, cudaEventRecord(p->ev_ready, caller));
    HR_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, p->ev_ready, 0));
    cudaStream_t st = ctx->comm_stream;
    HR_NCCL(ctx, N.GroupStart());
    for (int i = 0; i < n; i++)
        for (int r = 0; r < ctx->world; r++)
        {
            int rb, re;
            hr_shard_rows(items[i].H, r, ctx->world, &rb, &re);
            const bool last_band = re >= items[i].H;
            if (items[i].shift) { rb <<= items[i].shift; re <<= items[i].shift; }
            if (items[i].div > 1) { rb /= items[i].div; re = (re + items[i].div - 1) / items[i].div; }
            if (last_band || re > items[i].rows) re = items[i].rows;
            if (re <= rb) continue;
            char* p = static_cast<char*>(items[i].base) + (size_t)rb * items[i].row_bytes;
            HR_NCCL(ctx, N.Broadcast(p, p, (size_t)(re - rb) * items[i].row_bytes, ncclUint8, r, (ncclComm_t)ctx->nccl_comm, st));
        }
    HR_NCCL(ctx, N.GroupEnd());
    HR_CUDA(ctx, cudaEventRecord(p->ev_done, st));
    p->xchg_pending = true;
    return HR_OK;
}

extern "C" {

int hr_shard_unique_id(void* out_128_bytes)
{
    if (!out_128_bytes) return HR_ERR_INVALID_ARG;
    NcclApi& N = nccl();
    if (!N.ok) { hr_set_error(nullptr, "hr_shard_unique_id: libnccl.so.2 could not be loaded"); return HR_ERR_NCCL; }
    ncclUniqueId id;
    if (N.GetUniqueId(&id) != ncclSuccess) { hr_set_error(nullptr, "ncclGetUniqueId failed"); return HR_ERR_NCCL; }
    memcpy(out_128_bytes, &id, sizeof(id));
    return HR_OK;
}

int hr_shard_init(hr_ctx* ctx, int rank, int world, const void* unique_id_128_bytes)
{
    HR_REQUIRE(ctx, ctx && world >= 1 && rank >= 0 && rank < world, HR_ERR_INVALID_ARG, "hr_shard_init: bad rank/world");
    ctx->rank  = rank;
    ctx->world = world;
    if (world == 1) return HR_OK;
    HR_REQUIRE(ctx, unique_id_128_bytes, HR_ERR_INVALID_ARG, "hr_shard_init: unique id required for world > 1");
    NcclApi& N = nccl();
    HR_REQUIRE(ctx, N.ok, HR_ERR_NCCL, "hr_shard_init: libnccl.so.2 could not be loaded");
    HR_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, unique_id_128_bytes, sizeof(id));
    ncclComm_t comm;
    HR_NCCL(ctx, N.CommInitRank(&comm, world, id, rank));
    ctx->nccl_comm = comm;
    return HR_OK;
}

int hr_shard_shutdown(hr_ctx* ctx)
{
    if (!ctx) return HR_ERR_INVALID_ARG;
    if (ctx->nccl_comm) { nccl().CommDestroy((ncclComm_t)ctx->nccl_comm); ctx->nccl_comm = nullptr; }
    ctx->rank  = 0;
    ctx->world = 1;
    return HR_OK;
}

} // extern "C"
