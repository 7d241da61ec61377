// stats.cu — measurement helpers behind hr_pass_get_stats / hr_pass_output_checksum (not on the frame path).
//   * tile statistics: how many 8x8 tiles of the rows a render covered were on the denoise list (tile flag = 1) — bench.py
//     computes the a-trous roofline on the bytes the launch really processed (the reference's K2/K4 tile lists
//     `ray_traced_shadows.cpp:224` are this flag image here);
//   * ray counters: drained from the per-pass counter slots the trace kernels add to (traverse.cuh::count_rays);
//   * image checksum: order-independent 64-bit sum of mix(index, value) over the texels of an output — lets a sharded run prove
//     on the device that its gathered frame equals the single-GPU frame.
#include "hr_internal.h"

namespace {

__global__ void k_tile_stats(const uint8_t* __restrict__ flags, int TW, int t0, int t1, unsigned long long* __restrict__ out)
{
    unsigned long long n = 0;
    const size_t first = (size_t)t0 * TW, count = (size_t)(t1 - t0) * TW;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) n += flags[first + i] ? 1u : 0u;
    n = __reduce_add_sync(0xFFFFFFFFu, (unsigned)n);
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(out, n);
}

__global__ void k_drain_ray_counters(unsigned long long* __restrict__ ctr, unsigned long long* __restrict__ out)
{
    const int kind = threadIdx.x >> 5, slot = threadIdx.x & 31; // 64 threads: 2 kinds x 32 slots
    unsigned long long* p = ctr + ((size_t)kind * HR_RAY_CTR_SLOTS + slot) * HR_RAY_CTR_STRIDE;
    unsigned long long  v = *p;
    *p                    = 0ull;
    // 64-bit warp sum through two 32-bit halves
    const unsigned lo = __reduce_add_sync(0xFFFFFFFFu, (unsigned)(v & 0xFFFFu)), m1 = __reduce_add_sync(0xFFFFFFFFu, (unsigned)((v >> 16) & 0xFFFFu)),
                   m2 = __reduce_add_sync(0xFFFFFFFFu, (unsigned)((v >> 32) & 0xFFFFu)), hi = __reduce_add_sync(0xFFFFFFFFu, (unsigned)(v >> 48));
    if (slot == 0) out[kind] = (unsigned long long)lo + ((unsigned long long)m1 << 16) + ((unsigned long long)m2 << 32) + ((unsigned long long)hi << 48);
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long z)
{ // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// words: the image as 32-bit words (every format here has rows that are multiples of 2 bytes; the byte count is passed and the
// tail handled bytewise); word index is relative to the start of the IMAGE (not of the row range), so band sums add up.
__global__ void k_checksum(const uint8_t* __restrict__ base, size_t byte0, size_t byte1, unsigned long long* __restrict__ out)
{
    unsigned long long acc = 0;
    const size_t w0 = (byte0 + 3) / 4, w1 = byte1 / 4;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(base);
    for (size_t i = w0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < w1; i += (size_t)gridDim.x * blockDim.x)
        acc += mix64((i << 32) ^ (unsigned long long)w[i] ^ 0x9E3779B97F4A7C15ull);
    if (blockIdx.x == 0 && threadIdx.x == 0)
    { // unaligned head / tail bytes (byte ranges of whole rows of 1- or 2-byte texels)
        for (size_t b = byte0; b < byte1 && b < w0 * 4; b++) acc += mix64(((b | (1ull << 40)) << 8) ^ base[b]);
        for (size_t b = (w1 * 4 > byte0 ? w1 * 4 : byte0); b < byte1; b++) acc += mix64(((b | (1ull << 40)) << 8) ^ base[b]);
    }
    // block reduction
    __shared__ unsigned long long s[32];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32)
    {
        unsigned long long v = threadIdx.x < (blockDim.x >> 5) ? s[threadIdx.x] : 0ull;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        if (threadIdx.x == 0) atomicAdd(out, v);
    }
}

} // namespace

// out[0] = tiles with flag != 0 in tile rows [t0, t1)
void launch_tile_stats(const uint8_t* flags, int TW, int t0, int t1, unsigned long long* d_out, cudaStream_t st)
{
    if (t1 <= t0) return;
    k_tile_stats<<<64, 256, 0, st>>>(flags, TW, t0, t1, d_out);
}

// out[0] = primary, out[1] = secondary rays since the last drain (counters reset)
void launch_drain_ray_counters(unsigned long long* ctr, unsigned long long* d_out, cudaStream_t st) { k_drain_ray_counters<<<1, 64, 0, st>>>(ctr, d_out); }

void launch_checksum(const void* base, size_t byte0, size_t byte1, unsigned long long* d_out, cudaStream_t st)
{
    if (byte1 <= byte0) return;
    k_checksum<<<296, 256, 0, st>>>(static_cast<const uint8_t*>(base), byte0, byte1, d_out);
}
