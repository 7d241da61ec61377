// shard.cu — screen-space row-band sharding across the GPUs of one box (no reference counterpart: the
// reference is a single-GPU sample).
//
// One process per GPU.  A pass image of height H is split into `world` contiguous bands of whole 8-row tiles
// (hr_shard_rows).  Every rank keeps a full replica of the G-buffer, the BVH and the blue-noise tables, computes all
// stages on its band plus a recompute halo (rays and stencils are pure functions of replicated inputs, so the halo is
// exact).  Two things cross GPUs:
//   * PEER HISTORY (shadows / AO): next frame's reprojection may read ANY row of the history images under camera motion.
//     Instead of all-gathering ~16 B/pixel of history every frame, each rank keeps only its band and maps every peer's
//     history images into its address space (CUDA IPC over NVLink / NVSwitch); the temporal kernel fetches a history
//     texel from the GPU that owns its row (HistPeers) — the exchange is fused into the kernel and moves exactly the
//     texels the reprojection touches (a few boundary rows for a static camera).  Ordering is a per-pass frame tick:
//     after its last history write of frame N a rank stores N into every peer's tick array (st.release.sys); before
//     the reprojection of frame N+1 a one-warp kernel spins (ld.acquire.sys, 10 s time-out) until every peer's tick
//     is >= N.  History images are double buffered by frame parity, so that wait also covers the write-after-read hazard.
//   * the pass's FINAL OUTPUT (optional, hr_shard_set_gather; on by default): one NCCL group of per-band broadcasts (an
//     all-gather with unequal counts: 2160 rows = 270 tiles do not divide evenly by 8) on a side stream, so every rank
//     ends up with the complete denoised frame.  The reflections / DDGI passes still exchange their history this way.
// The 1/2/4/8-GPU results are bit-identical to the single-GPU result (tests/test_gpu_multi.py).
//
// NCCL is dlopen'ed on first use so the single-GPU library has no NCCL dependency; inside a PyTorch process the
// already-loaded torch-bundled libnccl.so.2 is picked up.
#include "hr_internal.h"
#include <dlfcn.h>
#include <nccl.h>
#include <vector>

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*)                                                                 = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int)                                          = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t)                                                                    = nullptr;
    ncclResult_t (*GroupStart)()                                                                               = nullptr;
    ncclResult_t (*GroupEnd)()                                                                                 = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t)       = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t)            = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t)                   = nullptr; // optional
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t)                         = nullptr; // optional
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
    LOAD(AllGather, "ncclAllGather");
    LOAD(Send, "ncclSend");
    LOAD(Recv, "ncclRecv");
    LOAD(GetErrorString, "ncclGetErrorString");
#undef LOAD
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.GroupStart && api.GroupEnd && api.Broadcast && api.AllGather && api.GetErrorString;
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

// How the final-output gather moves the bands: 0 = one ncclBroadcast per band, 1 = point-to-point (every rank sends its band to
// each peer and receives theirs: 7 independent NVSwitch transfers per rank instead of 8 serial broadcast rings).  hr_debug_set key 11.
int g_hr_gather_impl = 1;

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
            char*        p     = static_cast<char*>(items[i].base) + (size_t)rb * items[i].row_bytes;
            const size_t bytes = (size_t)(re - rb) * items[i].row_bytes;
            if (g_hr_gather_impl == 1 && N.Send && N.Recv)
            {
                if (r != ctx->rank) HR_NCCL(ctx, N.Recv(p, bytes, ncclUint8, r, (ncclComm_t)ctx->nccl_comm, st));
                else
                    for (int q = 0; q < ctx->world; q++)
                        if (q != ctx->rank) HR_NCCL(ctx, N.Send(p, bytes, ncclUint8, q, (ncclComm_t)ctx->nccl_comm, st));
            }
            else HR_NCCL(ctx, N.Broadcast(p, p, bytes, ncclUint8, r, (ncclComm_t)ctx->nccl_comm, st));
        }
    HR_NCCL(ctx, N.GroupEnd());
    HR_CUDA(ctx, cudaEventRecord(p->ev_done, st));
    p->xchg_pending = true;
    return HR_OK;
}


int hr_shard_exchange_rows(hr_pass* p, const RowRangeItem* items, int n, const int* row0, const int* row1, cudaStream_t st)
{
    hr_ctx* ctx = p->ctx;
    if (ctx->world <= 1 || !ctx->nccl_comm) return HR_OK;
    NcclApi& N = nccl();
    HR_NCCL(ctx, N.GroupStart());
    for (int i = 0; i < n; i++)
        for (int r = 0; r < ctx->world; r++)
        {
            if (row1[r] <= row0[r]) continue;
            char* ptr = static_cast<char*>(items[i].base) + (size_t)row0[r] * items[i].row_bytes;
            HR_NCCL(ctx, N.Broadcast(ptr, ptr, (size_t)(row1[r] - row0[r]) * items[i].row_bytes, ncclUint8, r, (ncclComm_t)ctx->nccl_comm, st));
        }
    HR_NCCL(ctx, N.GroupEnd());
    return HR_OK;
}

// ---- peer history -----------------------------------------------------------------------------------------------------
static int rt_init_bounds(hr_pass* p);

namespace {

struct TickPtrs { int* p[HR_MAX_RANKS]; };

__global__ void k_peer_signal(TickPtrs peers, int world, int self, int slot, int tick)
{
    const int r = threadIdx.x;
    if (r >= world || r == self || !peers.p[r]) return;
    __threadfence_system(); // the writes of the kernels before us in stream order become visible system-wide first
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(peers.p[r] + slot * HR_MAX_RANKS + self), "r"(tick) : "memory");
}

// ---- cooperative ray trace: push of the traced share + cost-balanced partition ----
struct CostPtrs { uint32_t* p[HR_MAX_RANKS]; };
struct MaskPtrs { uint32_t* p[HR_MAX_RANKS]; };

// After the ray-trace kernel: copy this rank's share of the mask image (rows [bounds[self], bounds[self+1])) into every
// peer's mask image with wide stores, drain the cost accumulation into every rank's cost table, reset the job counter,
// and — from the last block to finish — publish the ray-trace tick to the peers.
__global__ void __launch_bounds__(256) k_rt_push(MaskPtrs masks, CostPtrs costs, int* __restrict__ ctl, uint32_t* __restrict__ acc, TickPtrs ticks, int world, int self,
                                                 int parity, int MH, int MW, int tick)
{
    const int    b0 = ctl[self], b1 = ctl[self + 1];
    const size_t first = (size_t)b0 * MW, n = (size_t)(b1 - b0) * MW;
    const uint32_t* src = masks.p[self] + first;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
    if ((MW & 3) == 0)
    { // rows are 16-byte multiples and the images 256-byte aligned
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
        for (size_t i = tid; i < n / 4; i += nthr)
        {
            const uint4 v = s4[i];
            for (int r = 0; r < world; r++)
                if (r != self) reinterpret_cast<uint4*>(masks.p[r] + first)[i] = v;
        }
    }
    else
        for (size_t i = tid; i < n; i += nthr)
        {
            const uint32_t v = src[i];
            for (int r = 0; r < world; r++)
                if (r != self) masks.p[r][first + i] = v;
        }
    if (blockIdx.x == 0)
        for (int row = b0 + threadIdx.x; row < b1; row += blockDim.x)
        {
            const uint32_t v = acc[row];
            acc[row]         = 0u;
            for (int r = 0; r < world; r++) costs.p[r][(size_t)parity * MH + row] = v;
        }
    __threadfence_system();
    __syncthreads();
    __shared__ int s_last;
    if (threadIdx.x == 0)
    {
        const int done = atomicAdd(ctl + HR_MAX_RANKS + 2, 1);
        s_last         = done == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x == 0)
    {
        ctl[HR_MAX_RANKS + 2] = 0; // blocks-done counter
    }
    const int r = threadIdx.x;
    if (r < world && r != self && ticks.p[r])
    {
        __threadfence_system(); // all blocks' stores (ordered before their atomicAdd) are visible before the tick
        asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(ticks.p[r] + 1 * HR_MAX_RANKS + self), "r"(tick) : "memory");
    }
}

// One warp.  Lanes spin until every peer's ray-trace tick reached `rt_tick` and its history tick `hist_tick` (10 s time-out),
// then the warp computes next frame's partition: bounds[k] = first mask row whose cost prefix reaches k/world of the total.
// Every rank computes the same table from the same (complete) cost table.
__global__ void k_rt_wait_partition(const int* __restrict__ ticks, int rt_tick, int hist_tick, int* err, const uint32_t* __restrict__ cost, int MH, int world, int self,
                                    int cap, int* __restrict__ bounds)
{
    __shared__ unsigned long long s_prefix[1024 + 1];
    const int lane = threadIdx.x;
    if (lane < world && lane != self)
    {
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (;;)
        {
            int v, h = hist_tick;
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(ticks + HR_MAX_RANKS + lane) : "memory");
            if (hist_tick > 0) asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(h) : "l"(ticks + lane) : "memory");
            if (v >= rt_tick && h >= hist_tick) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 10000000000ull) { *err = 1 + lane; break; }
            __nanosleep(100);
        }
    }
    __syncwarp();
    const int          chunk = (MH + 31) / 32;
    unsigned long long local = 0;
    for (int i = lane * chunk; i < min((lane + 1) * chunk, MH); i++) local += __ldcg(cost + i);
    unsigned long long incl = local;
    for (int o = 1; o < 32; o <<= 1)
    {
        const unsigned long long v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += v;
    }
    const unsigned long long total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    unsigned long long       run   = incl - local;
    if (MH <= 1024)
    {
        for (int i = lane * chunk; i < min((lane + 1) * chunk, MH); i++) { run += __ldcg(cost + i); s_prefix[i + 1] = run; }
        if (lane == 0) s_prefix[0] = 0;
    }
    __syncwarp();
    if (lane == 0)
    {
        int prev  = 0;
        bounds[0] = 0;
        for (int k = 1; k < world; k++)
        {
            int b;
            if (total == 0 || MH > 1024) b = (int)(((long long)MH * k) / world); // no cost information: uniform split
            else
            {
                const unsigned long long target = (total * (unsigned long long)k) / (unsigned long long)world;
                int lo = prev, hi = MH; // first row index b with prefix[b] >= target
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_prefix[mid] >= target) hi = mid; else lo = mid + 1; }
                b = lo;
            }
            b = max(b, prev + 1);               // at least one row per rank
            b = min(b, prev + cap);             // the fixed launch grid covers at most `cap` rows
            b = min(b, MH - (world - k));       // leave a row for each remaining rank
            b = max(b, MH - (world - k) * cap); // the remaining ranks must be able to cover the rest
            bounds[k] = b;
            prev      = b;
        }
        bounds[world] = MH;
    }
}

__global__ void k_peer_wait(const int* __restrict__ ticks, int world, int self, int tick, int* err)
{
    const int r = threadIdx.x;
    if (r >= world || r == self) return;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;)
    {
        int v;
        asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(ticks + r) : "memory");
        if (v >= tick) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 10000000000ull) { *err = 1 + r; break; } // a peer died or the ranks render different frame counts: do not hang the GPU
        __nanosleep(200);
    }
}

// Reflections (interleaved cooperative ray trace): block b copies this rank's chunk c = self + b * world (8 image rows of RGBA16F
// texels) to every rank whose denoise stages read those rows; the last block to finish publishes the ray-trace tick.
#define PUSH_SPLIT 4
struct RowPtrs { uint2* p[HR_MAX_RANKS]; };
struct NeedRows { int r0[HR_MAX_RANKS], r1[HR_MAX_RANKS]; };
__global__ void __launch_bounds__(256) k_rt_push_chunks(RowPtrs imgs, NeedRows need, TickPtrs ticks, int* __restrict__ ctl, int world, int self, int W, int H, int tick)
{
    // PUSH_SPLIT blocks share one chunk (a chunk is 8 rows x W texels = 245 KB at 4K: one block would walk it in 60 dependent round trips)
    const int c = self + ((int)blockIdx.x / PUSH_SPLIT) * world, part = (int)blockIdx.x % PUSH_SPLIT, r0 = c * 8, r1 = min(r0 + 8, H);
    if (r0 < H)
        for (int q = 0; q < world; q++)
        {
            if (q == self) continue;
            const int a = max(r0, need.r0[q]), b = min(r1, need.r1[q]);
            if (b <= a) continue;
            const size_t first = (size_t)a * W, n = (size_t)(b - a) * W; // W is even: 16-byte pairs
            const uint4* s4 = reinterpret_cast<const uint4*>(imgs.p[self] + first);
            uint4*       d4 = reinterpret_cast<uint4*>(imgs.p[q] + first);
            for (size_t i = (size_t)part * blockDim.x + threadIdx.x; i < n / 2; i += (size_t)PUSH_SPLIT * blockDim.x) d4[i] = s4[i];
        }
    __threadfence_system();
    __syncthreads();
    __shared__ int s_last;
    if (threadIdx.x == 0)
    {
        const int done = atomicAdd(ctl + HR_MAX_RANKS + 2, 1);
        s_last         = done == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x == 0) ctl[HR_MAX_RANKS + 2] = 0;
    const int r = threadIdx.x;
    if (r < world && r != self && ticks.p[r])
    {
        __threadfence_system();
        asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(ticks.p[r] + 1 * HR_MAX_RANKS + self), "r"(tick) : "memory");
    }
}

int peer_alloc_ticks(hr_pass* p)
{
    hr_ctx* ctx = p->ctx;
    if (p->sync_ticks) return HR_OK;
    HR_CUDA(ctx, cudaMalloc((void**)&p->sync_ticks, sizeof(int) * 2 * HR_MAX_RANKS));
    HR_CUDA(ctx, cudaMemset(p->sync_ticks, 0, sizeof(int) * 2 * HR_MAX_RANKS));
    HR_CUDA(ctx, cudaHostAlloc((void**)&p->sync_error, sizeof(int), cudaHostAllocMapped));
    *p->sync_error = 0;
    return HR_OK;
}

} // namespace

void hr_peer_register(hr_pass* p, void* const* imgs, int n)
{
    p->n_hist = n;
    for (int k = 0; k < n; k++) p->hist_local[k] = imgs[k];
}

// Collective over the NCCL communicator: all-gather the CUDA IPC handles of this pass's history images and tick array,
// map every peer's.  Called from the first sharded render of the pass (all ranks render their passes in the same order).
int hr_peer_link_ipc(hr_pass* p, cudaStream_t st)
{
    hr_ctx* ctx = p->ctx;
    if (p->peers_linked || ctx->world <= 1 || !ctx->nccl_comm || p->n_hist == 0) return HR_OK;
    NcclApi& N = nccl();
    int      rc = peer_alloc_ticks(p);
    if (rc != HR_OK) return rc;
    const int    nh = p->n_hist + 1, world = ctx->world;
    const size_t hb = sizeof(cudaIpcMemHandle_t), mine = hb * nh;
    std::vector<cudaIpcMemHandle_t> h(nh), all((size_t)nh * world);
    for (int k = 0; k < p->n_hist; k++) HR_CUDA(ctx, cudaIpcGetMemHandle(&h[k], p->hist_local[k]));
    HR_CUDA(ctx, cudaIpcGetMemHandle(&h[p->n_hist], p->sync_ticks));
    char *d_send = nullptr, *d_recv = nullptr;
    HR_CUDA(ctx, cudaMalloc((void**)&d_send, mine));
    HR_CUDA(ctx, cudaMalloc((void**)&d_recv, mine * world));
    HR_CUDA(ctx, cudaMemcpyAsync(d_send, h.data(), mine, cudaMemcpyHostToDevice, st));
    HR_NCCL(ctx, N.AllGather(d_send, d_recv, mine, ncclUint8, (ncclComm_t)ctx->nccl_comm, st));
    HR_CUDA(ctx, cudaMemcpyAsync(all.data(), d_recv, mine * world, cudaMemcpyDeviceToHost, st));
    HR_CUDA(ctx, cudaStreamSynchronize(st));
    cudaFree(d_send);
    cudaFree(d_recv);
    for (int r = 0; r < world; r++)
    {
        if (r == ctx->rank)
        {
            for (int k = 0; k < p->n_hist; k++) p->hist_peer[r][k] = p->hist_local[k];
            p->peer_ticks[r] = p->sync_ticks;
            continue;
        }
        for (int k = 0; k < nh; k++)
        {
            void* m = nullptr;
            HR_CUDA(ctx, cudaIpcOpenMemHandle(&m, all[(size_t)r * nh + k], cudaIpcMemLazyEnablePeerAccess));
            if (k < p->n_hist) p->hist_peer[r][k] = m;
            else p->peer_ticks[r] = (int*)m;
        }
    }
    if (p->rt_bounds) { rc = rt_init_bounds(p); if (rc != HR_OK) return rc; }
    p->peers_linked = true;
    p->peers_ipc    = true;
    return HR_OK;
}

void hr_peer_unlink(hr_pass* p)
{
    if (p->peers_ipc)
        for (int r = 0; r < HR_MAX_RANKS; r++)
        {
            if (r == p->ctx->rank) continue;
            for (int k = 0; k < p->n_hist; k++)
                if (p->hist_peer[r][k]) cudaIpcCloseMemHandle(p->hist_peer[r][k]);
            if (p->peer_ticks[r]) cudaIpcCloseMemHandle(p->peer_ticks[r]);
        }
    p->peers_linked = p->peers_ipc = false;
    if (p->sync_ticks) cudaFree(p->sync_ticks);
    if (p->sync_error) cudaFreeHost(p->sync_error);
    p->sync_ticks = nullptr;
    p->sync_error = nullptr;
}

// History table for the reprojection kernel: images `img_k` / `aux_k` of every rank, rows split like hr_shard_rows(H).
void hr_peer_hist(const hr_pass* p, int img_k, int aux_k, int H, bool no_history, HistPeers* out)
{
    const hr_ctx* ctx = p->ctx;
    HistPeers     hp {};
    hp.no_history = no_history ? 1 : 0;
    if (p->peers_linked && ctx->world > 1)
    {
        hp.world = ctx->world;
        hp.self  = ctx->rank;
        for (int r = 0; r < ctx->world; r++)
        {
            int b, e;
            hr_shard_rows(H, r, ctx->world, &b, &e);
            hp.img[r]      = p->hist_peer[r][img_k];
            hp.aux[r]      = p->hist_peer[r][aux_k];
            hp.band_end[r] = e;
        }
    }
    else
    { // single GPU, or band emulation where the caller moves the bands itself (hr_shard_config + hr_pass_upload)
        hp.world       = 1;
        hp.self        = 0;
        hp.img[0]      = p->hist_local[img_k];
        hp.aux[0]      = p->hist_local[aux_k];
        hp.band_end[0] = H;
    }
    *out = hp;
}

int hr_peer_wait(hr_pass* p, int which, int tick, cudaStream_t st)
{
    if (!p->peers_linked || p->ctx->world <= 1 || tick <= 0) return HR_OK;
    hr_ctx* ctx = p->ctx;
    if (p->sync_error && *p->sync_error)
    {
        hr_set_error(ctx, "peer history: timed out waiting for rank %d's frame tick (peer stopped, or ranks rendered different frame counts)", *p->sync_error - 1);
        return HR_ERR_NCCL;
    }
    int* d_err = nullptr;
    HR_CUDA(ctx, cudaHostGetDevicePointer((void**)&d_err, p->sync_error, 0));
    k_peer_wait<<<1, 32, 0, st>>>(p->sync_ticks + which * HR_MAX_RANKS, ctx->world, ctx->rank, tick, d_err);
    ctx->launches++;
    return HR_OK;
}

int hr_peer_signal(hr_pass* p, int which, int tick, cudaStream_t st)
{
    if (!p->peers_linked || p->ctx->world <= 1) return HR_OK;
    TickPtrs t {};
    for (int r = 0; r < p->ctx->world; r++) t.p[r] = p->peer_ticks[r];
    k_peer_signal<<<1, 32, 0, st>>>(t, p->ctx->world, p->ctx->rank, which, tick);
    p->ctx->launches++;
    return HR_OK;
}

int hr_refl_push_chunks(hr_pass* p, int parity, int tick, const int* need0, const int* need1, cudaStream_t st)
{
    hr_ctx* ctx = p->ctx;
    RowPtrs  im {};
    NeedRows nd {};
    TickPtrs t {};
    for (int r = 0; r < ctx->world; r++)
    {
        im.p[r]  = static_cast<uint2*>(p->hist_peer[r][4 + parity]);
        nd.r0[r] = need0[r];
        nd.r1[r] = need1[r];
        t.p[r]   = p->peer_ticks[r];
    }
    const int n_chunks = (p->H + 7) / 8, mine = (n_chunks - ctx->rank + ctx->world - 1) / ctx->world;
    k_rt_push_chunks<<<(mine > 0 ? mine : 1) * PUSH_SPLIT, 256, 0, st>>>(im, nd, t, p->rt_bounds, ctx->world, ctx->rank, p->W, p->H, tick);
    ctx->launches++;
    return HR_OK;
}

// ---- shared ray mask + cost-balanced partition ------------------------------------------------------------------------
static int rt_init_bounds(hr_pass* p)
{ // first frame: uniform split of the mask rows
    hr_ctx*   ctx = p->ctx;
    const int MH = (p->H + 3) / 4, world = ctx->world;
    int       h[HR_MAX_RANKS + 1];
    for (int k = 0; k <= world; k++) h[k] = (int)(((long long)MH * k) / world);
    HR_CUDA(ctx, cudaMemcpy(p->rt_bounds, h, sizeof(int) * (world + 1), cudaMemcpyHostToDevice));
    return HR_OK;
}

int g_hr_force_shared_rt = 0; // hr_debug_set(4, 1): run the cooperative ray-trace kernel on a single GPU (overhead A/B)

bool hr_rt_share(hr_pass* p, int parity, RtShare* out)
{
    hr_ctx* ctx = p->ctx;
    const bool forced = g_hr_force_shared_rt && ctx->world == 1 && p->n_hist >= 7;
    if (forced && !p->hist_peer[0][6])
    {
        for (int k = 0; k < p->n_hist; k++) p->hist_peer[0][k] = p->hist_local[k];
        if (rt_init_bounds(p) != HR_OK) return false;
    }
    if (!forced && (!p->peers_linked || ctx->world <= 1 || p->n_hist < 7)) return false;
    RtShare sh {};
    sh.mask_local = static_cast<uint32_t*>(p->hist_local[4 + parity]);
    sh.bounds     = p->rt_bounds;
    sh.cost_acc   = p->rt_cost_acc;
    sh.world      = ctx->world;
    sh.self       = ctx->rank;
    *out          = sh;
    return true;
}

int hr_rt_share_finish(hr_pass* p, int parity, int tick, cudaStream_t st)
{
    hr_ctx*   ctx = p->ctx;
    const int MH = (p->H + 3) / 4, MW = (p->W + 7) / 8;
    CostPtrs  c {};
    MaskPtrs  m {};
    TickPtrs  t {};
    for (int r = 0; r < ctx->world; r++)
    {
        c.p[r] = static_cast<uint32_t*>(p->hist_peer[r][6]);
        m.p[r] = static_cast<uint32_t*>(p->hist_peer[r][4 + parity]);
        t.p[r] = p->peer_ticks[r];
    }
    k_rt_push<<<ctx->world > 1 ? 32 : 1, 256, 0, st>>>(m, c, p->rt_bounds, p->rt_cost_acc, t, ctx->world, ctx->rank, parity, MH, MW, tick);
    ctx->launches++;
    return HR_OK;
}

int hr_rt_wait_partition(hr_pass* p, int parity, int rt_tick, int hist_tick, cudaStream_t st)
{
    hr_ctx*   ctx = p->ctx;
    const int MH  = (p->H + 3) / 4;
    int*      d_err = nullptr;
    if (ctx->world > 1)
    {
        if (p->sync_error && *p->sync_error)
        {
            hr_set_error(ctx, "peer sync: timed out waiting for rank %d's frame tick (peer stopped, or ranks rendered different frame counts)", *p->sync_error - 1);
            return HR_ERR_NCCL;
        }
        HR_CUDA(ctx, cudaHostGetDevicePointer((void**)&d_err, p->sync_error, 0));
    }
    k_rt_wait_partition<<<1, 32, 0, st>>>(p->sync_ticks, rt_tick, hist_tick, d_err, p->rt_cost_all + (size_t)parity * MH, MH, ctx->world, ctx->rank,
                                         hr_rt_share_cap(MH, ctx->world), p->rt_bounds);
    ctx->launches++;
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
    HR_REQUIRE(ctx, world <= HR_MAX_RANKS, HR_ERR_UNSUPPORTED, "hr_shard_init: at most 8 ranks — one process per GPU of ONE box (peer history uses CUDA IPC mappings)");
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

// Same-process peers (N ranks emulated on one GPU, or a multi-GPU single-process host): rank `rank`'s history lives in
// `peer`'s images.  Every pass of the group must be linked with every other before the first sharded render.
int hr_shard_link_local(hr_pass* pass, int rank, hr_pass* peer)
{
    if (!pass || !peer) return HR_ERR_INVALID_ARG;
    hr_ctx* ctx = pass->ctx;
    HR_REQUIRE(ctx, rank >= 0 && rank < ctx->world && rank < HR_MAX_RANKS && pass->kind == peer->kind && pass->n_hist == peer->n_hist && pass->n_hist > 0 &&
                        pass->W == peer->W && pass->H == peer->H && !pass->peers_ipc,
               HR_ERR_INVALID_ARG, "hr_shard_link_local: passes do not match (kind / size / rank) or the pass has no peer history");
    int rc = peer_alloc_ticks(pass);
    if (rc != HR_OK) return rc;
    rc = peer_alloc_ticks(peer);
    if (rc != HR_OK) return rc;
    for (int k = 0; k < pass->n_hist; k++) pass->hist_peer[rank][k] = peer->hist_local[k];
    pass->peer_ticks[rank] = peer->sync_ticks;
    for (int k = 0; k < pass->n_hist; k++) pass->hist_peer[ctx->rank][k] = pass->hist_local[k];
    pass->peer_ticks[ctx->rank] = pass->sync_ticks;
    bool all = true;
    for (int r = 0; r < ctx->world; r++) all = all && pass->peer_ticks[r] != nullptr;
    pass->peers_linked = all;
    if (all && pass->rt_bounds) return rt_init_bounds(pass);
    return HR_OK;
}

int hr_shard_set_gather(hr_ctx* ctx, int gather_final_output)
{
    if (!ctx) return HR_ERR_INVALID_ARG;
    ctx->gather_final = gather_final_output != 0;
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
