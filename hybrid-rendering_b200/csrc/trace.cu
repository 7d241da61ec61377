// trace.cu — software BVH traversal + the two mask-producing ray-trace kernels.
//   K1  shadows/shadows_ray_trace.comp:89-132   (1 soft-shadow ray / pixel -> 1 bit)
//   K7  ao/ao_ray_trace.comp:90-126             (1 cosine-lobe AO ray / pixel -> 1 bit)
// Ray queries replace rayQueryEXT (ray_query.glsl:6-59): any-hit, t in (t_min, t_max), opaque, no culling.
//
// BUILD NOTE: this file is compiled with -fmad=false: the mask chain (det_math.cuh + ray_triangle below) is specified
// without implicit FMA contraction so that the packed visibility mask is bit-exact against the CPU oracle.  The slab
// test uses explicit fmaf (it only has to be conservative, the boxes are padded at build time).
//
// Mapping: one warp = one 8x4 pixel block = one mask word (bit = lane = (y&3)*8 + (x&7), exactly
// gl_LocalInvocationIndex of the reference's 8x4 workgroup), so the word is a single __ballot_sync.
#include "det_math.cuh"
#include "hr_internal.h"

namespace {

using det::V3;

struct Ray { V3 o, d; float tmin, tmax; };

// Moeller-Trumbore, fixed operation order (see oracle/orc_scene.h::ray_triangle for the CPU statement).
__device__ __forceinline__ bool ray_triangle(const float4 A, const float4 B, const float4 C, const Ray& r, float& t, float& u, float& v)
{
    const V3    v0 = det::mk(A.x, A.y, A.z), e1 = det::mk(B.x, B.y, B.z), e2 = det::mk(C.x, C.y, C.z);
    const V3    p   = det::cross(r.d, e2);
    const float dt  = det::dot(e1, p);
    if (dt == 0.0f) return false;
    const float inv = 1.0f / dt;
    const V3    tv  = det::sub(r.o, v0);
    u               = det::dot(tv, p) * inv;
    if (!(u >= 0.0f && u <= 1.0f)) return false;
    const V3 q = det::cross(tv, e1);
    v          = det::dot(r.d, q) * inv;
    if (!(v >= 0.0f && u + v <= 1.0f)) return false;
    t = det::dot(e2, q) * inv;
    return t > r.tmin && t < r.tmax;
}

#define STACK_SIZE 64
#define SENTINEL 0x7FFFFFFF

struct SlabSetup { float idx, idy, idz, ox, oy, oz; };
__device__ __forceinline__ SlabSetup slab_setup(const Ray& r)
{
    // |d| < 1e-18 is replaced by +-1e-18 for the BOX test only: with an infinite reciprocal the fma form
    // lo*inf - o*inf turns into NaN/-inf and would reject boxes the ray is inside of (not conservative).
    SlabSetup s;
    s.idx = 1.0f / (fabsf(r.d.x) > 1e-18f ? r.d.x : copysignf(1e-18f, r.d.x));
    s.idy = 1.0f / (fabsf(r.d.y) > 1e-18f ? r.d.y : copysignf(1e-18f, r.d.y));
    s.idz = 1.0f / (fabsf(r.d.z) > 1e-18f ? r.d.z : copysignf(1e-18f, r.d.z));
    s.ox  = r.o.x * s.idx;
    s.oy  = r.o.y * s.idy;
    s.oz  = r.o.z * s.idz;
    return s;
}

// Tests both children of a node; returns entry distances. fminf/fmaxf drop NaNs (0*inf) => conservative.
__device__ __forceinline__ void node_test(const float4* __restrict__ nodes, int node, const SlabSetup& s, float tmin, float tmax, bool& h0, bool& h1,
                                          float& tn0, float& tn1, int& c0, int& c1)
{
    const float4 n0 = __ldg(nodes + 4ull * node + 0);
    const float4 n1 = __ldg(nodes + 4ull * node + 1);
    const float4 nz = __ldg(nodes + 4ull * node + 2);
    const float4 ch = __ldg(nodes + 4ull * node + 3);
    c0 = __float_as_int(ch.x);
    c1 = __float_as_int(ch.y);
    float ax = fmaf(n0.x, s.idx, -s.ox), bx = fmaf(n0.y, s.idx, -s.ox);
    float ay = fmaf(n0.z, s.idy, -s.oy), by = fmaf(n0.w, s.idy, -s.oy);
    float az = fmaf(nz.x, s.idz, -s.oz), bz = fmaf(nz.y, s.idz, -s.oz);
    tn0      = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), tmin));
    float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tmax));
    h0       = tn0 <= tf;
    ax = fmaf(n1.x, s.idx, -s.ox); bx = fmaf(n1.y, s.idx, -s.ox);
    ay = fmaf(n1.z, s.idy, -s.oy); by = fmaf(n1.w, s.idy, -s.oy);
    az = fmaf(nz.z, s.idz, -s.oz); bz = fmaf(nz.w, s.idz, -s.oz);
    tn1 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), tmin));
    tf  = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tmax));
    h1  = tn1 <= tf;
}

// Any-hit traversal (while-while). Returns true as soon as one triangle is hit in (tmin, tmax).
__device__ bool trace_any(const BvhDev& bvh, const Ray& r)
{
    int       stack[STACK_SIZE];
    int       sp = 0;
    stack[sp++]  = SENTINEL;
    int             node = 0;
    const SlabSetup s    = slab_setup(r);
    while (node != SENTINEL)
    {
        while (node >= 0 && node != SENTINEL)
        {
            bool  h0, h1;
            float t0, t1;
            int   c0, c1;
            node_test(bvh.nodes, node, s, r.tmin, r.tmax, h0, h1, t0, t1, c0, c1);
            if (!h0 && !h1) node = stack[--sp];
            else
            {
                node = h0 ? c0 : c1;
                if (h0 && h1)
                {
                    if (t1 < t0) { int tmp = c1; c1 = node; node = tmp; }
                    if (sp < STACK_SIZE) stack[sp++] = c1;
                }
            }
        }
        if (node < 0)
        {
            const int leaf  = ~node;
            const int first = leaf >> 3, cnt = (leaf & 7) + 1;
            for (int k = 0; k < cnt; k++)
            {
                const float4 A = __ldg(bvh.tris + 3ull * (first + k));
                const float4 B = __ldg(bvh.tris + 3ull * (first + k) + 1);
                const float4 C = __ldg(bvh.tris + 3ull * (first + k) + 2);
                float        t, u, v;
                if (ray_triangle(A, B, C, r, t, u, v)) return true;
            }
            node = stack[--sp];
        }
    }
    return false;
}

// Closest hit; ties broken by the lowest primitive index (order independent).
__device__ bool trace_closest(const BvhDev& bvh, const Ray& r, float& best_t, uint32_t& best_prim, float& best_u, float& best_v)
{
    int stack[STACK_SIZE];
    int sp      = 0;
    stack[sp++] = SENTINEL;
    int node    = 0;
    best_t      = r.tmax;
    best_prim   = 0xFFFFFFFFu;
    best_u = best_v = 0.0f;
    const SlabSetup s = slab_setup(r);
    while (node != SENTINEL)
    {
        while (node >= 0 && node != SENTINEL)
        {
            bool  h0, h1;
            float t0, t1;
            int   c0, c1;
            node_test(bvh.nodes, node, s, r.tmin, best_t, h0, h1, t0, t1, c0, c1);
            if (!h0 && !h1) node = stack[--sp];
            else
            {
                node = h0 ? c0 : c1;
                if (h0 && h1)
                {
                    if (t1 < t0) { int tmp = c1; c1 = node; node = tmp; }
                    if (sp < STACK_SIZE) stack[sp++] = c1;
                }
            }
        }
        if (node < 0)
        {
            const int leaf  = ~node;
            const int first = leaf >> 3, cnt = (leaf & 7) + 1;
            for (int k = 0; k < cnt; k++)
            {
                const float4 A = __ldg(bvh.tris + 3ull * (first + k));
                const float4 B = __ldg(bvh.tris + 3ull * (first + k) + 1);
                const float4 C = __ldg(bvh.tris + 3ull * (first + k) + 2);
                float        t, u, v;
                if (ray_triangle(A, B, C, r, t, u, v))
                {
                    const uint32_t prim = __float_as_uint(A.w);
                    if (t < best_t || (t == best_t && prim < best_prim)) { best_t = t; best_prim = prim; best_u = u; best_v = v; }
                }
            }
            node = stack[--sp];
        }
    }
    return best_prim != 0xFFFFFFFFu;
}

__device__ __forceinline__ float2 load_oct_normal(const uint2* gb2, size_t idx)
{
    const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(gb2 + idx)); // first two halves = oct normal
    const __half2  h = *reinterpret_cast<const __half2*>(&w);
    return __half22float2(h);
}

// MODE 0 = shadows (K1), 1 = AO (K7).  Block = 256 threads = 8 warps; warp w covers mask word (bx*4 + (w&3), by*2 + (w>>2)).
template <int MODE>
__global__ void __launch_bounds__(256) k_ray_trace_mask(GBufLevelDev g, BvhDev bvh, FrameConsts fc, float p0, float p1, const uint8_t* __restrict__ sobol,
                                                         const uint8_t* __restrict__ sr, uint32_t* __restrict__ mask, int mrow0, int mrow1)
{
    const int MW   = (g.W + 7) >> 3;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mx = blockIdx.x * 4 + (warp & 3), my = mrow0 + blockIdx.y * 2 + (warp >> 2);
    if (mx >= MW || my >= mrow1) return; // whole warp exits together
    const int x = mx * 8 + (lane & 7), y = my * 4 + (lane >> 3);
    uint32_t  result = 0;
    if (x < g.W && y < g.H)
    {
        const size_t idx   = (size_t)y * g.W + x;
        const float  depth = __ldg(g.depth + idx);
        if (depth != 1.0f)
        {
            const float  u = ((float)x + 0.5f) / (float)g.W, v = ((float)y + 0.5f) / (float)g.H;
            const V3     P  = det::world_position_from_depth(u, v, depth, fc.view_proj_inverse);
            const float2 e  = load_oct_normal(g.gb2, idx);
            const V3     N  = det::octohedral_to_direction(e.x, e.y);
            const float  r0 = det::sample_blue_noise(x, y, (int)fc.num_frames, 0, sobol, sr);
            const float  r1 = det::sample_blue_noise(x, y, (int)fc.num_frames, 1, sobol, sr);
            Ray          r;
            r.tmin = 0.01f;
            if (MODE == 0)
            {
                r.o = det::add(P, det::scale(N, p0)); // bias
                float att;
                det::fetch_light_properties_shadow(fc.light, P, N, r0, r1, r.d, r.tmax, att);
                if (att > 0.0f) result = trace_any(bvh, r) ? 0u : 1u;
            }
            else
            {
                r.o    = det::add(P, det::scale(N, p1)); // bias
                r.d    = det::sample_cosine_lobe(N, r0, r1);
                r.tmax = p0; // ray_length
                result = trace_any(bvh, r) ? 0u : 1u;
            }
        }
    }
    const uint32_t word = __ballot_sync(0xFFFFFFFFu, result != 0);
    if (lane == 0) mask[(size_t)my * MW + mx] = word;
}

__global__ void k_trace_any(BvhDev bvh, const float* __restrict__ rays, size_t n, uint32_t* __restrict__ out)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = rays + 8 * i;
    Ray          r;
    r.o = det::mk(p[0], p[1], p[2]); r.tmin = p[3];
    r.d = det::mk(p[4], p[5], p[6]); r.tmax = p[7];
    out[i] = trace_any(bvh, r) ? 1u : 0u;
}

__global__ void k_trace_closest(BvhDev bvh, const float* __restrict__ rays, size_t n, float* __restrict__ out_t, uint32_t* __restrict__ out_prim,
                                float* __restrict__ out_uv)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = rays + 8 * i;
    Ray          r;
    r.o = det::mk(p[0], p[1], p[2]); r.tmin = p[3];
    r.d = det::mk(p[4], p[5], p[6]); r.tmax = p[7];
    float    t, u, v;
    uint32_t prim;
    trace_closest(bvh, r, t, prim, u, v);
    out_t[i]    = t;
    out_prim[i] = prim;
    if (out_uv) { out_uv[2 * i] = u; out_uv[2 * i + 1] = v; }
}

} // namespace

static inline dim3 mask_grid(int W, int mrow0, int mrow1) { return dim3(((W + 7) / 8 + 3) / 4, (mrow1 - mrow0 + 1) / 2, 1); }

void launch_shadows_ray_trace(const GBufLevelDev& g, const BvhDev& bvh, const FrameConsts& fc, float bias, const uint8_t* sobol, const uint8_t* sr,
                              uint32_t* mask, int row0, int row1, cudaStream_t st)
{
    const int mrow0 = row0 / 4, mrow1 = (row1 + 3) / 4;
    if (mrow1 <= mrow0) return;
    k_ray_trace_mask<0><<<mask_grid(g.W, mrow0, mrow1), 256, 0, st>>>(g, bvh, fc, bias, 0.0f, sobol, sr, mask, mrow0, mrow1);
}

void launch_ao_ray_trace(const GBufLevelDev& g, const BvhDev& bvh, const FrameConsts& fc, float ray_length, float bias, const uint8_t* sobol,
                         const uint8_t* sr, uint32_t* mask, int row0, int row1, cudaStream_t st)
{
    const int mrow0 = row0 / 4, mrow1 = (row1 + 3) / 4;
    if (mrow1 <= mrow0) return;
    k_ray_trace_mask<1><<<mask_grid(g.W, mrow0, mrow1), 256, 0, st>>>(g, bvh, fc, ray_length, bias, sobol, sr, mask, mrow0, mrow1);
}

void launch_trace_any(const BvhDev& bvh, const float* rays, size_t n, uint32_t* out, cudaStream_t st)
{
    if (!n) return;
    k_trace_any<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(bvh, rays, n, out);
}

void launch_trace_closest(const BvhDev& bvh, const float* rays, size_t n, float* out_t, uint32_t* out_prim, float* out_uv, cudaStream_t st)
{
    if (!n) return;
    k_trace_closest<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(bvh, rays, n, out_t, out_prim, out_uv);
}
