// bvh_build.cu — device LBVH build (replaces the driver BLAS/TLAS build,
// external/dwSampleFramework/extras/ray_traced_scene.cpp:196-248 + src/mesh.cpp:169-231; the reference rebuilds its
// TLAS every frame, src/main.cpp:74).
//
// Pipeline (all on the caller's stream, deterministic):
//   1. k_tri_bounds   per-triangle AABB + scene AABB (ordered-int atomics)
//   2. k_morton       63-bit Morton code of the AABB centre (21 bits / axis)
//   3. cub radix sort (key = Morton, value = primitive index; stable => ties keep primitive order)
//   4. k_hierarchy    Karras 2012 radix tree: children, leaf ranges, parents
//   5. k_fit          bottom-up AABB fit (one atomic flag per internal node)
//   6. k_pack         traversal layout: 64-byte nodes holding both child boxes; sub-trees with <= LEAF_MAX
//                     triangles are collapsed into one leaf (their triangles are contiguous in Morton order)
//
// Node layout (4 x float4):  n0 = (c0.lo.x, c0.hi.x, c0.lo.y, c0.hi.y)   n1 = (c1.lo.x, c1.hi.x, c1.lo.y, c1.hi.y)
//                            nz = (c0.lo.z, c0.hi.z, c1.lo.z, c1.hi.z)   ch = (int c0, int c1, -, -) as bits
// child >= 0: internal node index;  child < 0: leaf, ~child = (first_tri << 3) | (count - 1).
// Triangle layout (3 x float4, leaf order): (v0.xyz, prim bits) (e1.xyz, 0) (e2.xyz, 0), e = v - v0.
// Boxes are padded by 2^-16 of the scene extent so the slab test is conservative w.r.t. the fp32 ray/triangle test.
#include "hr_internal.h"
#include <cub/device/device_radix_sort.cuh>
#include <cfloat>

#define LEAF_MAX 4

namespace {

__device__ __forceinline__ int   f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7FFFFFFF; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

__global__ void k_init_bounds(int* b)
{
    if (threadIdx.x < 3) b[threadIdx.x] = f2ord(FLT_MAX);
    else if (threadIdx.x < 6) b[threadIdx.x] = f2ord(-FLT_MAX);
}

__global__ void k_tri_bounds(const float* __restrict__ verts, uint32_t n, float* __restrict__ aabb, int* __restrict__ bounds)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float    lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    if (i < n)
    {
        const float* p = verts + 9ull * i;
        for (int a = 0; a < 3; a++)
        {
            lo[a] = fminf(fminf(p[a], p[3 + a]), p[6 + a]);
            hi[a] = fmaxf(fmaxf(p[a], p[3 + a]), p[6 + a]);
            aabb[6ull * i + a]     = lo[a];
            aabb[6ull * i + 3 + a] = hi[a];
        }
    }
    for (int a = 0; a < 3; a++)
    {
        for (int o = 16; o; o >>= 1)
        {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xFFFFFFFFu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xFFFFFFFFu, hi[a], o));
        }
        if ((threadIdx.x & 31) == 0)
        {
            atomicMin(bounds + a, f2ord(lo[a]));
            atomicMax(bounds + 3 + a, f2ord(hi[a]));
        }
    }
}

__device__ __forceinline__ uint64_t spread21(uint32_t v)
{
    uint64_t x = v & 0x1FFFFFull;
    x          = (x | x << 32) & 0x1F00000000FFFFull;
    x          = (x | x << 16) & 0x1F0000FF0000FFull;
    x          = (x | x << 8) & 0x100F00F00F00F00Full;
    x          = (x | x << 4) & 0x10C30C30C30C30C3ull;
    x          = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void k_morton(const float* __restrict__ aabb, uint32_t n, const int* __restrict__ bounds, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t q[3];
    for (int a = 0; a < 3; a++)
    {
        float lo = ord2f(bounds[a]), hi = ord2f(bounds[3 + a]);
        float c  = 0.5f * (aabb[6ull * i + a] + aabb[6ull * i + 3 + a]);
        float e  = hi - lo;
        float t  = e > 0.0f ? (c - lo) / e : 0.0f;
        t        = fminf(fmaxf(t, 0.0f), 1.0f);
        q[a]     = (uint32_t)fminf(t * 2097152.0f, 2097151.0f);
    }
    keys[i] = (spread21(q[0]) << 2) | (spread21(q[1]) << 1) | spread21(q[2]);
    vals[i] = i;
}

// delta(i,j) = length of the common prefix of key i and key j (ties broken by index), -1 outside [0,n)
__device__ __forceinline__ int delta(const uint64_t* __restrict__ keys, int n, int i, int j)
{
    if (j < 0 || j >= n) return -1;
    uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz((uint32_t)i ^ (uint32_t)j);
    return __clzll((long long)(a ^ b));
}

// Karras 2012, "Maximizing parallelism in the construction of BVHs, octrees, and k-d trees", Alg. on p.4.
// Internal nodes 0..n-2; leaves are referred to as n-1+k in the parent array.
__global__ void k_hierarchy(const uint64_t* __restrict__ keys, int n, int2* __restrict__ children, int2* __restrict__ ranges, int* __restrict__ parent)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d       = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin    = delta(keys, n, i, i - d);
    int lmax    = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j     = i + l * d;
    int dnode = delta(keys, n, i, j);
    int s     = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1)
    {
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    int gamma = i + s * d + min(d, 0);
    int first = min(i, j), last = max(i, j);
    int left  = (first == gamma) ? (n - 1 + gamma) : gamma;          // leaf ids offset by n-1
    int right = (last == gamma + 1) ? (n - 1 + gamma + 1) : gamma + 1;
    children[i] = make_int2(left, right);
    ranges[i]   = make_int2(first, last);
    parent[left]  = i;
    parent[right] = i;
    if (i == 0) parent[0] = -1;
}

__global__ void k_fit(const float* __restrict__ tri_aabb, const uint32_t* __restrict__ sorted_prim, int n, const int2* __restrict__ children,
                      const int* __restrict__ parent, float* __restrict__ node_aabb, int* __restrict__ flags)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int node = parent[n - 1 + k];
    while (node >= 0)
    {
        if (atomicAdd(flags + node, 1) == 0) return; // first arrival: the sibling subtree is not finished yet
        __threadfence();
        int2  c = children[node];
        float lo[3], hi[3];
        for (int s = 0; s < 2; s++)
        {
            int          ch = s ? c.y : c.x;
            const float* b  = ch >= n - 1 ? tri_aabb + 6ull * sorted_prim[ch - (n - 1)] : node_aabb + 6ull * ch;
            for (int a = 0; a < 3; a++)
            {
                float l = __ldcg(b + a), h = __ldcg(b + 3 + a);
                lo[a] = s ? fminf(lo[a], l) : l;
                hi[a] = s ? fmaxf(hi[a], h) : h;
            }
        }
        for (int a = 0; a < 3; a++)
        {
            __stcg(node_aabb + 6ull * node + a, lo[a]);
            __stcg(node_aabb + 6ull * node + 3 + a, hi[a]);
        }
        __threadfence();
        node = parent[node];
    }
}

__device__ __forceinline__ int encode_child(int ch, int n, const int2* __restrict__ ranges)
{
    if (ch >= n - 1) return ~(((ch - (n - 1)) << 3) | 0);
    int2 r = ranges[ch];
    int  cnt = r.y - r.x + 1;
    if (cnt <= LEAF_MAX) return ~((r.x << 3) | (cnt - 1));
    return ch;
}

__global__ void k_pack_nodes(int n, const int2* __restrict__ children, const int2* __restrict__ ranges, const float* __restrict__ tri_aabb,
                             const uint32_t* __restrict__ sorted_prim, const float* __restrict__ node_aabb, const int* __restrict__ bounds,
                             float4* __restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    float ext = 0.0f;
    for (int a = 0; a < 3; a++) ext = fmaxf(ext, ord2f(bounds[3 + a]) - ord2f(bounds[a]));
    const float pad = ext * (1.0f / 65536.0f) + 1e-7f;
    int2  c = children[i];
    float b[2][6];
    for (int s = 0; s < 2; s++)
    {
        int          ch  = s ? c.y : c.x;
        const float* src = ch >= n - 1 ? tri_aabb + 6ull * sorted_prim[ch - (n - 1)] : node_aabb + 6ull * ch;
        for (int a = 0; a < 3; a++)
        {
            b[s][a]     = src[a] - pad;
            b[s][3 + a] = src[3 + a] + pad;
        }
    }
    out[4ull * i + 0] = make_float4(b[0][0], b[0][3], b[0][1], b[0][4]);
    out[4ull * i + 1] = make_float4(b[1][0], b[1][3], b[1][1], b[1][4]);
    out[4ull * i + 2] = make_float4(b[0][2], b[0][5], b[1][2], b[1][5]);
    out[4ull * i + 3] = make_float4(__int_as_float(encode_child(c.x, n, ranges)), __int_as_float(encode_child(c.y, n, ranges)), 0.0f, 0.0f);
}

// n <= LEAF_MAX (or n == 1): a single node whose child 0 is the leaf [0,n) with the scene box, child 1 is empty.
__global__ void k_pack_tiny(int n, const int* __restrict__ bounds, float4* __restrict__ out)
{
    float ext = 0.0f, lo[3], hi[3];
    for (int a = 0; a < 3; a++)
    {
        lo[a] = ord2f(bounds[a]);
        hi[a] = ord2f(bounds[3 + a]);
        ext   = fmaxf(ext, hi[a] - lo[a]);
    }
    const float pad = ext * (1.0f / 65536.0f) + 1e-7f;
    out[0] = make_float4(lo[0] - pad, hi[0] + pad, lo[1] - pad, hi[1] + pad);
    // child 1 is unused: a degenerate box far outside any scene (an inverted box would read as infinite in a min/max slab
    // test) that still refers to a valid leaf, so even a visit is harmless.
    out[1] = make_float4(1e30f, 1e30f, 1e30f, 1e30f);
    out[2] = make_float4(lo[2] - pad, hi[2] + pad, 1e30f, 1e30f);
    out[3] = make_float4(__int_as_float(~((0 << 3) | (n - 1))), __int_as_float(~((0 << 3) | 0)), 0.0f, 0.0f);
}

__global__ void k_pack_tris(const float* __restrict__ verts, const uint32_t* __restrict__ sorted_prim, uint32_t n, float4* __restrict__ out)
{
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t     prim = sorted_prim[k];
    const float* p    = verts + 9ull * prim;
    out[3ull * k + 0] = make_float4(p[0], p[1], p[2], __uint_as_float(prim));
    out[3ull * k + 1] = make_float4(p[3] - p[0], p[4] - p[1], p[5] - p[2], 0.0f);
    out[3ull * k + 2] = make_float4(p[6] - p[0], p[7] - p[1], p[8] - p[2], 0.0f);
}

} // namespace

int hr_bvh_build(hr_scene* sc, cudaStream_t st)
{
    hr_ctx*        ctx = sc->ctx;
    const uint32_t n   = sc->n_tris;
    const int      T   = 256;
    const int      gb  = (int)((n + T - 1) / T);
    k_init_bounds<<<1, 32, 0, st>>>(sc->d_bounds_i);
    k_tri_bounds<<<gb, T, 0, st>>>(sc->d_tri_verts, n, sc->d_tri_aabb, sc->d_bounds_i);
    k_morton<<<gb, T, 0, st>>>(sc->d_tri_aabb, n, sc->d_bounds_i, sc->d_keys, sc->d_vals);
    HR_CHECK_LAUNCH(ctx);
    ctx->launches += 3;
    size_t need = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, need, sc->d_keys, sc->d_keys_sorted, sc->d_vals, sc->d_vals_sorted, (int)n, 0, 63, st);
    if (need > sc->sort_tmp_bytes)
    {
        if (sc->d_sort_tmp) cudaFree(sc->d_sort_tmp);
        HR_CUDA(ctx, cudaMalloc(&sc->d_sort_tmp, need));
        sc->sort_tmp_bytes = need;
    }
    HR_CUDA(ctx, cub::DeviceRadixSort::SortPairs(sc->d_sort_tmp, sc->sort_tmp_bytes, sc->d_keys, sc->d_keys_sorted, sc->d_vals, sc->d_vals_sorted,
                                                 (int)n, 0, 63, st));
    if (n > LEAF_MAX)
    {
        const int gi = (int)((n - 1 + T - 1) / T);
        HR_CUDA(ctx, cudaMemsetAsync(sc->d_flags, 0, sizeof(int) * (n - 1), st));
        k_hierarchy<<<gi, T, 0, st>>>(sc->d_keys_sorted, (int)n, sc->d_children, sc->d_ranges, sc->d_parent);
        k_fit<<<gb, T, 0, st>>>(sc->d_tri_aabb, sc->d_vals_sorted, (int)n, sc->d_children, sc->d_parent, sc->d_node_aabb, sc->d_flags);
        k_pack_nodes<<<gi, T, 0, st>>>((int)n, sc->d_children, sc->d_ranges, sc->d_tri_aabb, sc->d_vals_sorted, sc->d_node_aabb, sc->d_bounds_i, sc->d_nodes);
        sc->n_nodes = n - 1;
        ctx->launches += 3;
    }
    else
    {
        k_pack_tiny<<<1, 1, 0, st>>>((int)n, sc->d_bounds_i, sc->d_nodes);
        sc->n_nodes = 1;
        ctx->launches += 1;
    }
    k_pack_tris<<<gb, T, 0, st>>>(sc->d_tri_verts, sc->d_vals_sorted, n, sc->d_tris);
    ctx->launches += 1;
    HR_CHECK_LAUNCH(ctx);
    return HR_OK;
}

BvhDev hr_bvh_view(const hr_scene* sc)
{
    BvhDev v;
    v.nodes         = sc->d_nodes;
    v.tris          = sc->d_tris;
    v.root_is_valid = sc->n_tris > 0;
    return v;
}
